"""Run the UNMODIFIED reference (staged under baseline/_ref/src by baseline/stage_reference.py) on the host cores.

This is the CPU arm of the comparison (`bench.py --impl reference`, `cpu_baseline.kind == "reference"`): every rollout is
`agent.rollout_episode(PBO_Env(problem, <Algo>_Optimizer(config)))` of the reference's own classes -- src/tester.py:196-202 -- with its own
seeding convention `np.random.seed(run)` (tester.py:198) plus `torch.manual_seed(run)` (the reference leaves torch unseeded), one rollout
per job, jobs spread over a process pool (the same embarrassingly-parallel split the GPU path uses over lanes), one BLAS / torch thread
per worker.  Nothing of metabox_b200/ or oracle/ is imported here.

A *bounded* rollout (max_steps) is obtained WITHOUT touching reference code: `BoundedEnv` subclasses the reference's PBO_Env and reports
`done` after max_steps environment steps, so the agent's own `rollout_episode` loop ends early; everything it calls is unmodified."""
import argparse
import os
import sys
import time
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, '_ref', 'src')
STUBS = ['deap', 'deap.base', 'deap.creator', 'deap.tools', 'deap.cma', 'deap.algorithms', 'skopt', 'matplotlib', 'matplotlib.pyplot',
         'openpyxl', 'cmaes', 'gym', 'gym.spaces']        # imported at package-import time by out-of-scope baselines only
ALGOS = {  # algo -> (optimizer class, agent pickle relative to src/, steps are whole generations?)
    'rlepso': ('RLEPSO_Optimizer', 'agent_model/test/bbob_easy/RLEPSO_Agent.pkl'),
    'rlepso-protein': ('RLEPSO_Optimizer', 'agent_model/rollout/protein_easy/RLEPSO_Agent/checkpoint20.pkl'),
    'lde': ('LDE_Optimizer', 'agent_model/test/bbob_easy/LDE_Agent.pkl'),
    'deddqn': ('DE_DDQN_Optimizer', 'agent_model/test/bbob_easy/DE_DDQN_Agent.pkl'),
    'gleet': ('GLEET_Optimizer', 'agent_model/rollout/bbob_easy/GLEET_Agent/checkpoint20.pkl'),
}


def available():
    return os.path.exists(os.path.join(HERE, '_ref', 'MANIFEST.json'))


def activate(src=SRC):
    sys.dont_write_bytecode = True
    if not os.path.isdir(src):
        raise RuntimeError(f'staged reference not found at {src}: run `python baseline/stage_reference.py` in the build container')
    if src not in sys.path:
        sys.path.insert(0, src)
    for m in STUBS:
        sys.modules.setdefault(m, MagicMock())
    return src


def make_config(problem, dim, difficulty='easy'):
    """The constants src/config.py:74-102 derives, as the Namespace the reference's constructors expect (and mutate)."""
    c = argparse.Namespace(problem=problem, dim=dim, difficulty=difficulty, upperbound=5, device='cpu', train_batch_size=1, test_batch_size=1,
                           maxFEs=2000 * dim, n_logpoint=50, max_learning_step=1500000, n_checkpoint=20, agent_save_dir='/tmp/_mbx_ref_agents/')
    if problem in ('protein', 'protein-torch'):
        c.dim, c.maxFEs, c.n_logpoint = 12, 1000, 5
    c.log_interval = c.maxFEs // c.n_logpoint
    c.save_interval = c.max_learning_step // c.n_checkpoint
    return c


_STATE = {}


def _setup(algo, problem, dim, difficulty):
    """Per worker process: import the reference once, build the dataset and unpickle the shipped agent once."""
    key = (algo, problem, dim, difficulty)
    if key in _STATE:
        return _STATE[key]
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    activate()
    import copy
    import pickle
    import torch
    torch.set_num_threads(1)
    import optimizer as ropt
    import agent as ragent  # noqa: F401  (unpickling needs the agent classes)
    from environment import PBO_Env
    from utils import construct_problem_set
    cfg = make_config(problem, dim, difficulty)
    cwd = os.getcwd()
    os.chdir(SRC)                       # the protein loader opens its data directory relative to src/
    try:
        train, test = construct_problem_set(cfg)
    finally:
        os.chdir(cwd)
    problems = list(train.data) + list(test.data)
    opt_name, pkl = ALGOS[algo]
    with open(os.path.join(SRC, pkl), 'rb') as f:
        ag = pickle.load(f)
    opt = getattr(ropt, opt_name)(copy.deepcopy(cfg))

    class BoundedEnv(PBO_Env):
        def __init__(self, problem, optimizer, max_steps):
            super().__init__(problem, optimizer)
            self._left = max_steps

        def step(self, action):
            s, r, d = super().step(action)
            self._left -= 1
            return s, r, (d or self._left <= 0)

    _STATE[key] = (problems, ag, opt, PBO_Env, BoundedEnv)
    return _STATE[key]


def rollout_job(job):
    """job = (algo, problem, dim, difficulty, problem index into train+test, run seed, max_steps or None) -> (fes, seconds, final cost)."""
    algo, problem, dim, difficulty, k, run, max_steps = job
    problems, ag, opt, PBO_Env, BoundedEnv = _setup(algo, problem, dim, difficulty)
    import numpy as np
    import torch
    p = problems[k]
    t = time.perf_counter()
    np.random.seed(run)                 # tester.py:198
    torch.manual_seed(run)
    p.reset()                           # tester.py:199
    env = PBO_Env(p, opt) if max_steps is None else BoundedEnv(p, opt, max_steps)
    info = ag.rollout_episode(env)      # tester.py:202
    init_fes = 0
    return int(info['fes']) - init_fes, time.perf_counter() - t, float(info['cost'][-1])


_TRAIN = {}


def lde_train_job(job):
    """job = (problem, dim, difficulty, problem index, seed) -> (fes, seconds): ONE `LDE_Agent.train_episode(PBO_Env(problem, LDE_Optimizer))`
    of the unmodified reference (lde_agent.py:85-145: 20 trajectories x 50 generations at the reference's fixed NP = 50, then one REINFORCE
    step), with a freshly constructed agent per worker process (src/trainer.py:160-161 drives exactly this call per problem)."""
    problem, dim, difficulty, k, seed = job
    problems, _, opt, PBO_Env, _ = _setup('lde', problem, dim, difficulty)
    import copy
    import numpy as np
    import torch
    key = (problem, dim, difficulty)
    if key not in _TRAIN:
        import agent as ragent
        cfg = make_config(problem, dim, difficulty)
        cfg.agent_save_dir = f'/tmp/_mbx_ref_agents/{os.getpid()}/'
        torch.manual_seed(0)
        _TRAIN[key] = ragent.LDE_Agent(copy.deepcopy(cfg))
    ag = _TRAIN[key]
    p = problems[k]
    t = time.perf_counter()
    np.random.seed(seed)
    torch.manual_seed(seed)
    p.reset()
    env = PBO_Env(p, opt)
    ag.train_episode(env)
    # evaluations of the episode: every trajectory re-initialises the population (NP) and runs TRAJECTORY_LENGTH generations of NP
    NP, n_traj, n_len = 50, 20, 50
    return n_traj * (NP + n_len * NP), time.perf_counter() - t, 0.0


def problem_names(algo, problem, dim, difficulty):
    problems = _setup(algo, problem, dim, difficulty)[0]
    return [str(p) for p in problems]


def _warm(args):
    _setup(*args)
    return os.getpid()


def make_pool(procs, algo, problem, dim, difficulty):
    """Worker processes with the reference imported, the dataset built and the agent unpickled (outside any timed region)."""
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    pool = ctx.Pool(procs)
    pool.map(_warm, [(algo, problem, dim, difficulty)] * (4 * procs), chunksize=1)
    return pool


def timed_batch(pool, jobs, fn=None):
    """Run the jobs (longest expected first when the caller sorted them) -> (sum of fes, wall seconds, per-job (fes, seconds, cost))."""
    t = time.perf_counter()
    out = list(pool.imap(fn or rollout_job, jobs, chunksize=1))
    wall = time.perf_counter() - t
    return sum(o[0] for o in out), wall, out
