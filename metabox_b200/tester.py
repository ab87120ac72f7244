"""Batched counterparts of `Tester.test`, `test_for_random_search` and `rollout` (`src/tester.py:180-263, 266-352,
355-407`): the nested `for problem / for agent / for run / while not done` loops become ONE lane table per agent that
is swept on the GPU (sharded over ranks when torch.distributed is initialised), and the results are written in the
reference's own on-disk layout so that `logger.py` / the post-processing scripts read them unchanged:

    test.pkl                    {'cost': {problem: {agent: [runs x [51 floats]]}}, 'fes': {problem: {agent: [runs ints]}},
                                 'T0': ms, 'T1': {agent: ms}, 'T2': {agent: ms}}                       (tester.py:123-127)
    random_search_baseline.pkl  same layout over train + test with the single agent 'Random_search'   (tester.py:355-407)
    rollout.pkl                 {'cost' | 'fes' | 'return': {problem: {agent: [checkpoints x [5 runs]]}}} (tester.py:282-352)

Differences that are documented rather than hidden: the run seed of the reference (`np.random.seed(run)`) becomes the
Philox key (seed, global lane id), so individual curves differ from a numpy-RNG run while their distribution is the
same; T2 (optimisation time per run) is the amortised time of the batched sweep (sweep seconds / lanes), T1 the GPU
time to evaluate one run's maxFEs candidates of the first test problem (`cal_t1`), T0 the reference's CPU loop (`cal_t0`).
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np

N_RUNS_TEST, N_RUNS_ROLLOUT = 51, 5           # tester.py:121, :318
AGENT_ALGO = {'RLEPSO_Agent': 'rlepso', 'LDE_Agent': 'lde', 'DE_DDQN_Agent': 'deddqn', 'GLEET_Agent': 'gleet', 'RL_PSO_Agent': 'rl_pso',
              'QLPSO_Agent': 'qlpso', 'DEDQN_Agent': 'dedqn', 'RL_HPSDE_Agent': 'rl_hpsde', 'L2L_Agent': 'l2l', 'NRLPSO_Agent': 'nrlpso', 'LES_Agent': 'les'}


def cal_t0(dim, fes):
    """tester.py:61-76, verbatim semantics (a CPU calibration loop, milliseconds)."""
    T0 = 0
    for _ in range(10):
        start = time.perf_counter()
        for _ in range(fes):
            x = np.random.rand(dim)
            x + x
            x / (x + 2)
            x * x
            np.sqrt(x)
            np.log(x)
            np.exp(x)
        T0 += (time.perf_counter() - start) * 1000
    return T0 / 10


def construct_problem_set(config, data_folder=None):
    """utils.py:4-27."""
    from . import problem as mp
    name = config.problem[:-6] if config.problem.endswith('-torch') else config.problem
    if name in ('bbob', 'bbob-noisy'):
        return mp.BBOB_Dataset.get_datasets(name, config.dim, getattr(config, 'upperbound', 5),
                                            difficulty=getattr(config, 'difficulty', 'easy'))
    if name == 'protein':
        return mp.Protein_Docking_Dataset.get_datasets('protein', difficulty=getattr(config, 'difficulty', 'easy'),
                                                       data_folder=data_folder or getattr(config, 'protein_data', None))
    raise ValueError(name + ' is not defined!')


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _check_policy(algo, policy):
    """A policy object of the WRONG backbone (e.g. an RLEPSO actor handed to the NRLPSO sweep) used to fail deep inside torch with a dtype
    error; the known policy classes are checked against the backbone up front.  Unknown callables (user policies) pass."""
    from . import agent as ma
    known = {'rlepso': (ma.RLEPSO_Actor, ma.RLEPSO_FusedActor), 'lde': (ma.LDE_PolicyNet,), 'deddqn': (ma.DE_DDQN_Net, ma.DE_DDQN_FusedNet),
             'gleet': (ma.GLEET_Actor, ma.GLEET_FusedActor), 'rl_pso': (ma.RL_PSO_PolicyNet, ma.RL_PSO_FusedPolicy),
             'qlpso': (ma.QLPSO_QTable, ma.QLPSO_FusedQTable), 'nrlpso': (ma.QLPSO_QTable, ma.QLPSO_FusedQTable),
             'rl_hpsde': (ma.QLPSO_QTable, ma.QLPSO_FusedQTable), 'dedqn': (ma.DEDQN_Net,), 'l2l': (ma.L2L_PolicyNet,)}
    every = tuple({c for cs in known.values() for c in cs})
    if algo in known and isinstance(policy, every) and not isinstance(policy, known[algo]):
        raise ValueError(f'{type(policy).__name__} is not a policy of the {algo} backbone (expected one of '
                         f'{[c.__name__ for c in known[algo]]})')


def _make_sweep(algo, pack, lanes, policy, config, seed, lane_base, stripes):
    from . import sweep as ms
    n_log = config.n_logpoint
    _check_policy(algo, policy)

    def make(ln, lb):
        if algo == 'rlepso':
            return ms.RLEPSO_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        if algo == 'lde':
            return ms.LDE_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb, NP=policy.NP)
        if algo == 'gleet':                                   # policy: any torch callable state [B,NP,27] -> action [B,NP]
            return ms.GLEET_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        if algo == 'rl_pso':                                  # policy: state [B,2D] -> action [B]
            return ms.RL_PSO_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        if algo == 'les':                                     # policy: {'attn': [68], 'mlp': [178]} network parameters (the agent's best solution)
            return ms.LES_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        if algo == 'nrlpso':                                  # policy: state [B] int32 in 0..3 -> strategy [B] in 0..3 (tabular)
            return ms.NRLPSO_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        if algo == 'l2l':                                     # policy: agent.L2L_PolicyNet (float64 LSTM with projection)
            return ms.L2L_Sweep(pack, ln, policy, seed=seed, lane_base=lb)
        if algo == 'rl_hpsde':                                # policy: state [B] int32 in 0..3 -> action [B] in 0..3 (tabular)
            return ms.RL_HPSDE_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        if algo == 'dedqn':                                   # policy: landscape features [B,4] -> Q [B,3] (greedy argmax inside the sweep)
            return ms.DEDQN_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        if algo == 'qlpso':                                   # policy: q_state [B] int32 -> action [B] in 0..3
            return ms.QLPSO_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)
        return ms.DE_DDQN_Sweep(pack, ln, policy, config.maxFEs, n_log, seed=seed, lane_base=lb)

    return ms.StripedSweep(make, lanes, lane_base, stripes) if stripes > 1 and len(lanes) > stripes else make(lanes, lane_base)


def sweep_rows(algo, pack, n_problems, n_runs, policy, config, seed=0, stripes=2):
    """All (problem, run) rollouts of one agent -> (host rows [n_problems * n_runs, 53] in lane order, seconds).
    Lanes are sharded over the ranks of the default process group; ONE all-gather brings the rows together."""
    import torch
    from . import sweep as ms
    rank, world = _dist()
    inst, _ = ms.lane_table(n_problems, n_runs)
    lo, hi = ms.shard_range(len(inst), rank, world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sw = _make_sweep(algo, pack, inst[lo:hi], policy, config, seed, lo, stripes)
    rows = ms.gather_results(sw.run(), world, n_max=-(-len(inst) // world))
    torch.cuda.synchronize()
    return rows.cpu().numpy(), time.perf_counter() - t0


def rows_to_results(rows, problems, n_runs, agent_name, out):
    """rows [n_problems * n_runs, 53] -> out['cost'|'fes'(|'return')][problem][agent] (lists, like the reference)."""
    for k, p in enumerate(problems):
        blk = rows[k * n_runs:(k + 1) * n_runs]
        out['cost'].setdefault(str(p), {})[agent_name] = [[float(v) for v in r[:51]] for r in blk]
        out['fes'].setdefault(str(p), {})[agent_name] = [int(r[51]) for r in blk]
        if 'return' in out:
            out['return'].setdefault(str(p), {})[agent_name] = [float(r[52]) for r in blk]
    return out


class Tester:
    """`Tester(config, policies).test()` -- config fields used: problem, difficulty, dim, maxFEs, n_logpoint,
    log_interval, upperbound, test_log_dir, agent_for_cp (names out of RLEPSO_Agent / LDE_Agent / DE_DDQN_Agent / GLEET_Agent /
    RL_PSO_Agent / QLPSO_Agent),
    t_optimizer_for_cp (only 'Random_search' has a B200 path).  `policies`: {agent name: torch module on the GPU}.
    Every policy that has a single-kernel twin in the library is swapped for it (`agent.fuse`: same weights, one launch per step, validated
    against the torch modules and the reference's known answers in tests/test_gpu_z_policies.py); `config.fused_policies = False` keeps the
    torch modules."""

    def __init__(self, config, policies, device=None, data_folder=None):
        import torch
        from . import problem as mp
        self.config = config
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.train_set, self.test_set = construct_problem_set(config, data_folder)
        self.policies = dict(policies)
        if getattr(config, 'fused_policies', True):
            from .agent import fuse
            self.policies = {k: fuse(v) for k, v in self.policies.items()}
        for name in getattr(config, 'agent_for_cp', list(self.policies)):
            if name not in AGENT_ALGO:
                raise ValueError(f'{name}: no B200 backbone (supported: {sorted(AGENT_ALGO)})')
            if name not in self.policies:
                raise KeyError(f'policy for {name} missing')
        self.agent_names = list(getattr(config, 'agent_for_cp', list(self.policies)))
        self.t_optimizers = [o for o in getattr(config, 't_optimizer_for_cp', [])]
        for o in self.t_optimizers:
            if o != 'Random_search':
                raise ValueError(f'{o}: no B200 path (only Random_search among the classic optimizers)')
        self.pack = mp.InstancePack(self.test_set.data, self.device)
        self.log_dir = getattr(config, 'test_log_dir', None)
        self.test_results = {'cost': {}, 'fes': {}, 'T0': 0., 'T1': {}, 'T2': {}}

    def cal_t1(self):
        """tester.py:79-88 on the GPU: milliseconds to evaluate maxFEs candidates of the first test problem."""
        import torch
        if getattr(self, '_t1', None) is None:
            cfg, NP = self.config, 100
            L = max(1, cfg.maxFEs // NP)
            x = np.random.rand(L, NP, self.pack.dim)
            self.pack.eval_lanes(np.zeros(L, dtype=np.int32), x, noise_free=True)          # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                self.pack.eval_lanes(np.zeros(L, dtype=np.int32), x, noise_free=True)
            torch.cuda.synchronize()
            self._t1 = (time.perf_counter() - t0) * 1000 / 10
        return self._t1

    def _random_search(self, problems, pack, n_runs, seed):
        """Random_search over all (problem, run) lanes, sharded over the ranks like the agents' sweeps (one gather of the rows)."""
        import torch
        from . import optimizer as mo, sweep as ms
        rank, world = _dist()
        inst, _ = ms.lane_table(len(problems), n_runs)
        lo, hi = ms.shard_range(len(inst), rank, world)
        opt = mo.Random_search(self.config)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = opt.run_lanes(pack, inst[lo:hi], seed=seed, lane_base=lo)      # one launch per lane batch, whole episodes in-kernel
        rows = np.zeros((hi - lo, 53), dtype=np.float32)
        for b, r in enumerate(res):
            c = list(r['cost'])
            c += [c[-1]] * (51 - len(c))                       # tester.py:204-205
            rows[b, :51], rows[b, 51] = c[:51], r['fes']
        if world > 1:
            t = torch.as_tensor(rows).to(pack.device)
            rows = ms.gather_results({'curve': t[:, :51], 'fes': t[:, 51], 'ret': t[:, 52]}, world, n_max=-(-len(inst) // world)).cpu().numpy()
        return rows.astype(np.float64), time.perf_counter() - t0

    def test(self, n_runs=N_RUNS_TEST, seed=0, stripes=2, random_search_baseline=True):
        R = self.test_results
        cfg = self.config
        R['T0'] = cal_t0(cfg.dim, cfg.maxFEs)                  # tester.py:183
        probs = self.test_set.data
        for p in probs:
            R['cost'][str(p)], R['fes'][str(p)] = {}, {}
        for name in self.agent_names:
            rows, secs = sweep_rows(AGENT_ALGO[name], self.pack, len(probs), n_runs, self.policies[name], cfg, seed, stripes)
            rows_to_results(rows, probs, n_runs, name, R)
            R['T2'][name] = 1000. * secs / max(1, len(rows))   # amortised optimisation time per run (ms)
            R['T1'][name] = self.cal_t1()                      # time to evaluate one run's maxFEs candidates (ms)
        for o in self.t_optimizers:
            rows, secs = self._random_search(probs, self.pack, n_runs, seed)
            rows_to_results(rows, probs, n_runs, o, R)
            R['T1'][o] = R['T2'][o] = 1000. * secs / max(1, len(rows))
        rank, _ = _dist()
        out = {'test': R}
        if random_search_baseline:
            out['random_search_baseline'] = self.test_for_random_search(n_runs, seed)
        if self.log_dir and rank == 0:
            os.makedirs(self.log_dir, exist_ok=True)
            with open(os.path.join(self.log_dir, 'test.pkl'), 'wb') as f:
                pickle.dump(R, f, -1)
            if random_search_baseline:
                with open(os.path.join(self.log_dir, 'random_search_baseline.pkl'), 'wb') as f:
                    pickle.dump(out['random_search_baseline'], f, -1)
        return out

    def test_for_random_search(self, n_runs=N_RUNS_TEST, seed=0):
        """tester.py:355-407: Random_search over train + test."""
        from . import problem as mp
        entire = self.train_set.data + self.test_set.data
        pack = mp.InstancePack(entire, self.device)
        R = {'cost': {}, 'fes': {}, 'T0': cal_t0(self.config.dim, self.config.maxFEs), 'T1': {}, 'T2': {}}
        rows, secs = self._random_search(entire, pack, n_runs, seed + 1)
        rows_to_results(rows, entire, n_runs, 'Random_search', R)
        R['T1']['Random_search'] = R['T2']['Random_search'] = 1000. * secs / max(1, len(rows))
        return R


def rollout(config, checkpoints, n_runs=N_RUNS_ROLLOUT, seed=0, stripes=2, device=None, data_folder=None):
    """tester.py:266-352.  `checkpoints`: {agent name: [policy module of checkpoint 0, 1, ...]}.  Every (checkpoint,
    problem, run) of the TRAIN set is one lane; writes config.rollout_log_dir/rollout.pkl on rank 0."""
    import torch
    from . import problem as mp
    device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
    train_set, _ = construct_problem_set(config, data_folder)
    probs = train_set.data
    pack = mp.InstancePack(probs, device)
    R = {'cost': {}, 'fes': {}, 'return': {}}
    for name, nets in checkpoints.items():
        if name not in AGENT_ALGO:
            raise ValueError(f'{name}: no B200 backbone (supported: {sorted(AGENT_ALGO)})')
        for p in probs:
            for key in R:
                R[key].setdefault(str(p), {})[name] = []
        for ck, net in enumerate(nets):
            if getattr(config, 'fused_policies', True):
                from .agent import fuse
                net = fuse(net)
            rows, _ = sweep_rows(AGENT_ALGO[name], pack, len(probs), n_runs, net, config, seed + 7919 * ck, stripes)
            tmp = rows_to_results(rows, probs, n_runs, name, {'cost': {}, 'fes': {}, 'return': {}})
            for p in probs:
                for key in R:
                    R[key][str(p)][name].append(tmp[key][str(p)][name])
    rank, _ = _dist()
    log_dir = getattr(config, 'rollout_log_dir', None)
    if log_dir and rank == 0:
        os.makedirs(log_dir, exist_ok=True)
        with open(os.path.join(log_dir, 'rollout.pkl'), 'wb') as f:
            pickle.dump(R, f, -1)
    return R
