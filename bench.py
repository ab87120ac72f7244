#!/usr/bin/env python
"""bench.py -- candidate evaluations / second of the MetaBBO rollout hot path.

Workload (BASELINE.json metric "candidate evals/sec on BBOB dim=30 pop=100"): complete RLEPSO rollouts
(`agent.rollout_episode(PBO_Env(problem, RLEPSO_Optimizer))`) on all 24 BBOB functions, dim 30, pop 100,
maxFEs 60000, `--runs` independent runs per function per GPU (default 256 -> 6144 lanes per GPU, weak
scaling).  One bench *step* = one complete sweep of all lanes (init_population + every generation until
every lane is done + the final all-gather of cost curves).  evals are counted with the reference's `fes`
accounting.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--runs R]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (see the contract in DESIGN.md section 7).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM, MAXFES = 30, 60000
# workload -> (suite, difficulty/subset, NP, algorithmic bytes per evaluation, description)
WORKLOADS = {
    'rlepso': dict(suit='bbob', subset='all', NP=100, bytes_per_eval=24 * DIM + 24, kernel='rlepso_step_kernel<30,BBOB>',
                   desc='RLEPSO rollouts, BBOB 24 fns',
                   bound_note='instruction issue (62 % of issue slots busy, DRAM 4.5 % in the ncu capture profiles/r1e_ncu_rlepso_step_d30.txt), not HBM'),
    'lde': dict(suit='bbob-noisy', subset='all', NP=100, bytes_per_eval=8 * DIM + 20.4, kernel='lde_step_kernel<30,BBOB>',
                desc='LDE rollouts, BBOB-noisy 30 fns',
                bound_note='instruction issue (49 % of issue slots busy, FP64 pipe 16 %; profiles/r1e_ncu_lde_step_d30.txt), not HBM'),
    'deddqn': dict(suit='bbob', subset='difficult-test', NP=100, bytes_per_eval=32 * DIM + 404,
                   kernel='deddqn_step_kernel<30,BBOB>', desc='DE-DDQN rollouts, BBOB difficult test set 18 fns',
                   bound_note='latency of a chain of short dependent phases, one trial per lane per step (issue 34 % busy; profiles/r1g_ncu_deddqn.txt), not HBM'),
    'gleet': dict(suit='bbob', subset='all', NP=100, bytes_per_eval=24 * DIM + 204, kernel='gleet_step_kernel<30,BBOB>',
                  desc='GLEET rollouts, BBOB 24 fns',
                  bound_note='instruction issue and long-scoreboard stalls in the feature tail (issue 32 % busy; profiles/r1g_ncu_gleet.txt), not HBM'),
    # BASELINE config C5: the RLEPSO test sweep over the 280 protein-docking instances (28 complexes x 10 starts), 51 runs each, dim 12,
    # maxFEs 1000 -> 14 280 lanes of 10 generations; a FIXED job: N GPUs shard its lanes (strong scaling)
    'protein': dict(suit='protein', subset='all', NP=100, dim=12, maxFEs=1000, n_logpoint=5, runs=51, scaling='strong',
                    bytes_per_eval=24 * 12 + 24, kernel='rlepso_step_kernel<12,PROTEIN>',
                    desc='RLEPSO test sweep, protein docking 280 instances (config C5)',
                    bound_note='FP64 ALU of the pairwise energy (4.6e5 flop per evaluation against 312 B of state; profiles/r1g_ncu_rlepso-protein.txt), not HBM'),
    # BASELINE config C4: the LDE TRAIN rollout -- per learn step 20 trajectories x 50 generations on each of the 30 noisy functions (600
    # lanes) + one REINFORCE step + one gradient all-reduce (lde_agent.py:85-145); a FIXED job per learn step (strong scaling)
    'lde-train': dict(suit='bbob-noisy', subset='all', NP=100, scaling='strong', bytes_per_eval=8 * DIM + 20.4, kernel='lde_step_kernel<30,BBOB>',
                      desc='LDE train rollout + REINFORCE step, BBOB-noisy 30 fns x 20 trajectories x 50 generations per learn step (config C4)',
                      bound_note='launch latency of 50 dependent (policy, step) pairs on 600 lanes per learn step, then autograd; not HBM'),
}
METRIC = 'candidate evals/sec on BBOB dim=30 pop=100'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='rlepso', choices=sorted(WORKLOADS),
                    help='rlepso = the headline (BASELINE.json metric); lde / deddqn = the DE-family configs C4 / C3; '
                         'gleet = the GLEET backbone (SURVEY 8f N1), policy chosen by --gleet-policy')
    ap.add_argument('--gleet-policy', default='fused', choices=['actor', 'fused', 'stand-in'],
                    help="actor = the reference's attention actor with the weights of its shipped checkpoint 20 (agent.GLEET_Actor, torch); "
                         'fused = the same actor as ONE CUDA kernel per generation (agent.GLEET_FusedActor -> mbx_gleet_policy); '
                         'stand-in = a per-particle MLP 27->64->1 (what the r1g / r1h gleet lines under profiles/ were measured with)')
    ap.add_argument('--torch-policy', action='store_true',
                    help='keep the policy a torch module between the step kernels (round-1 path).  Default: the single-kernel twin of the '
                         'policy -- rlepso: agent.RLEPSO_FusedActor -> mbx_rlepso_policy; deddqn: Q-network + argmax, '
                         'agent.DE_DDQN_FusedNet -> mbx_deddqn_policy; gleet: agent.GLEET_FusedActor -> mbx_gleet_policy')
    ap.add_argument('--steps-per-graph', type=int, default=None,
                    help='environment steps captured per CUDA graph (default: 1 for the population backbones, 10 for deddqn); more steps = fewer '
                         'graph launches, finished lanes idle through the surplus steps')
    ap.add_argument('--runs', type=int, default=None, help='runs per function per GPU (default 256)')
    ap.add_argument('--cpu-gens', type=int, default=None, help='generations (deddqn: steps) per lane in one CPU sample step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-kind', default='auto', choices=['auto', 'reference', 'port'],
                    help='CPU arm: reference = the unmodified reference staged under baseline/_ref; port = the oracle restatement; '
                         'auto = reference when staged')
    ap.add_argument('--stripes', type=int, default=None,
                    help='independent lane stripes per GPU, each on its own CUDA stream + graph (default 3; deddqn 6)')
    ap.add_argument('--lane-order', default=None, choices=['instance-major', 'run-major', 'stripe-balanced'],
                    help='order of the (instance, run) lanes inside a GPU (default: instance-major; lde: stripe-balanced = every '
                         'stripe gets ~runs/stripes runs of every function, instance-major inside)')
    ap.add_argument('--poll', type=int, default=None, help='environment steps between two host polls of the done flags (default 64 generations; deddqn persistent: evaluations per launch, default 250)')
    ap.add_argument('--launches-per-poll', type=int, default=None, help='deddqn (persistent): launches issued back to back between two host polls of `done` (default 16)')
    ap.add_argument('--no-graph', action='store_true', help='launch policy + step kernels eagerly instead of replaying a CUDA graph')
    a = ap.parse_args()
    if a.runs is None:
        a.runs = WORKLOADS[a.workload].get('runs', 256)          # C3: 256 runs in lockstep; C5: 51 runs
    if a.stripes is None:
        a.stripes = 6 if (a.workload == 'deddqn' and a.torch_policy) else 3   # deddqn persistent: 576 CTAs on 444 resident slots -- 3 streams fill each other's last wave       # the 1-FE-per-step chain is latency bound: more chains in flight
    if a.lane_order is None:
        a.lane_order = 'stripe-balanced' if a.workload == 'lde' else 'instance-major'     # measured: profiles/, DESIGN.md section 6
    return a


# ---------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu, self.first = [], None, gpu_index, 0

    def mark(self):
        """Rows sampled from here on count (the sampler is started before the warm-up so that nvidia-smi's own start-up --
        NVML initialisation over all GPUs of the box -- does not fall into the timed region)."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        rows = self.rows[self.first:]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def policy_weights(workload, gleet_policy='actor'):
    """Shipped policies (extracted from the reference pickles into tests/golden/agents.npz / agents2.npz).  The shipped LDE policy
    only fits NP = 50 (lde_agent.py:37); the NP = 100 workload uses a randomly initialised policy of the right width."""
    p = os.path.join(ROOT, 'tests', 'golden', 'agents.npz')
    if workload == 'gleet':
        p2 = os.path.join(ROOT, 'tests', 'golden', 'agents2.npz')
        if gleet_policy == 'stand-in' or not os.path.exists(p2):
            return None, 'stand-in per-particle MLP 27->64->1 (random init, seed 0)'
        G = np.load(p2)
        return ({k[len('gleet/'):]: np.asarray(G[k]) for k in G.files if k.startswith('gleet/')},
                'GLEET attention actor, reference rollout checkpoint 20 (bbob_easy)')
    if workload == 'protein':
        G = np.load(os.path.join(ROOT, 'tests', 'golden', 'agents_protein.npz'))
        return ({k[len('rlepso/'):]: np.asarray(G[k]) for k in G.files if k.startswith('rlepso/')},
                'RLEPSO actor of the reference rollout checkpoint 20 (protein_easy)')
    if workload == 'lde' or not os.path.exists(p):
        return None, 'random init (seed 0)'
    G = np.load(p)
    pre = workload + '/'
    return {k[len(pre):]: np.asarray(G[k]) for k in G.files if k.startswith(pre)}, f'reference {workload} agent pickle (bbob_easy)'


def suite_fids(workload):
    W = WORKLOADS[workload]
    if W['suit'] == 'bbob-noisy':
        return list(range(101, 131))
    if W['subset'] == 'difficult-test':
        return [f for f in range(1, 25) if f not in (1, 5, 6, 10, 15, 20)]       # bbob.py:918, difficulty='difficult'
    return list(range(1, 25))


# ---------------------------------------------------------------------------------------------- reference arm
# what one CPU step is, per workload, when the UNMODIFIED reference runs (baseline/ref_runner.py): (ref algo, problem, difficulty, subset,
# max_steps per rollout (None = the complete rollout), note)
REF_RUNS = {
    'rlepso': ('rlepso', 'bbob', 'easy', 'all', None, 'complete rollouts (maxFEs 60000)'),
    'lde': ('lde', 'bbob-noisy', 'easy', 'all', None, "complete rollouts; the reference's LDE hard-codes NP = 50 (lde_optimizer.py:10, lde_agent.py:37), "
            'so its CPU figure is at pop 50 with the shipped policy'),
    'deddqn': ('deddqn', 'bbob', 'difficult', 'test', 3000, 'first 3000 one-evaluation steps of each rollout'),
    'gleet': ('gleet', 'bbob', 'easy', 'all', 150, 'first 150 generations of each rollout'),
    'protein': ('rlepso-protein', 'protein', 'easy', 'all', None, 'complete rollouts (maxFEs 1000)'),
}


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def run_reference_unmodified(args):
    """`--impl reference` proper: the reference's own classes, staged byte-identically under baseline/_ref (baseline/stage_reference.py),
    `agent.rollout_episode(PBO_Env(problem, <Algo>_Optimizer))` per (function, run) job on all host cores (src/tester.py:196-202)."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_runner as rr
    W = WORKLOADS[args.workload]
    algo, problem, difficulty, subset, max_steps, note = REF_RUNS[args.workload]
    if args.cpu_gens is not None:                                   # explicit bound: the first --cpu-gens environment steps of each rollout
        max_steps, note = args.cpu_gens, f'first {args.cpu_gens} environment steps of each rollout'
    cores = os.cpu_count() or 1
    dim = W.get('dim', DIM)
    pool = rr.make_pool(cores, algo, problem, dim, difficulty)
    names = rr.problem_names(algo, problem, dim, difficulty)       # train + test order of the reference's dataset factory
    n_train = {('bbob', 'easy'): 18, ('bbob', 'difficult'): 6, ('bbob-noisy', 'easy'): 22, ('bbob-noisy', 'difficult'): 8,
               ('protein', 'easy'): 200}[(problem, difficulty)]
    ks = list(range(n_train, len(names))) if subset == 'test' else list(range(len(names)))
    if len(ks) > 4 * cores:                                          # protein: 280 instances -> an evenly spaced sample of 4 per core
        ks = [ks[(j * len(ks)) // (4 * cores)] for j in range(4 * cores)]
        note += f' of an evenly spaced sample of {len(ks)} of the {len(names)} instances'
    runs = max(1, -(-cores // len(ks)))                              # >= 1 job per core per step
    # warm-up steps are short (20 environment steps per job): they fault the code in and give the per-function cost used to order the
    # timed jobs longest-first (a straggling Gallagher / Weierstrass rollout started last would leave the other cores idle)
    cost = {k: 1.0 for k in ks}
    for _ in range(max(1, args.warmup)):
        _, _, out = rr.timed_batch(pool, [(algo, problem, dim, difficulty, k, 0, 20) for k in ks])
        cost = {k: o[1] for k, o in zip(ks, out)}
    order = sorted(ks, key=lambda k: -cost[k])
    tot_fes, tot_s = 0, 0.0
    for step in range(args.steps):
        jobs = [(algo, problem, dim, difficulty, k, step * runs + r, max_steps) for k in order for r in range(runs)]
        fes, wall, _ = rr.timed_batch(pool, jobs)
        tot_fes += fes; tot_s += wall
    pool.close(); pool.join()
    NPref = 50 if args.workload == 'lde' else W['NP']
    sample = (f'{len(ks)} fns x {runs} run(s) per step, {note}; dim {dim}, pop {NPref}; {cores} worker processes (1 thread each); '
              f'warm-up steps are 20-step rollouts; glibc malloc tunables keep numpy temporaries off mmap')
    return tot_fes, tot_s, cores, 'reference', sample, f'shipped reference agent pickle ({rr.ALGOS[algo][1]})'


def run_reference_lde_train(args):
    """C4 on the CPU: `LDE_Agent.train_episode(PBO_Env(problem, LDE_Optimizer))` of the unmodified reference (one learn step per problem:
    20 trajectories x 50 generations + REINFORCE; lde_agent.py:85-145, driven per problem by trainer.py:160-161), one episode per job, a
    fresh agent per worker, all host cores.  The reference fixes NP = 50."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_runner as rr
    cores = os.cpu_count() or 1
    pool = rr.make_pool(cores, 'lde', 'bbob-noisy', DIM, 'easy')
    n = len(rr.problem_names('lde', 'bbob-noisy', DIM, 'easy'))
    per = max(1, -(-cores // n))
    jobs = [('bbob-noisy', DIM, 'easy', k, r) for r in range(per) for k in range(n)]
    rr.timed_batch(pool, jobs[:cores], fn=rr.lde_train_job)          # warm-up: agent construction, code faulted in
    tot_fes, tot_s = 0, 0.0
    for step in range(args.steps):
        fes, wall, _ = rr.timed_batch(pool, [(a, b, c, k, r + 1000 * step) for (a, b, c, k, r) in jobs], fn=rr.lde_train_job)
        tot_fes += fes; tot_s += wall
    pool.close(); pool.join()
    sample = (f'{len(jobs)} train episodes per step ({n} noisy fns x {per}), each 20 trajectories x 50 generations + one REINFORCE step at the '
              f"reference's fixed pop 50, dim {DIM}; {cores} worker processes (1 thread each)")
    return tot_fes, tot_s, cores, 'reference', sample, 'freshly initialised LDE_Agent per worker (reference constructor)'


def run_reference_port(args):
    """Fallback / second figure: the oracle port (`oracle/`, pinned against the reference by tests/golden) on all host cores, each step a
    bounded sample: every function x runs x `--cpu-gens` generations."""
    import torch
    from oracle import agents as oag, rollout as orr
    W = WORKLOADS[args.workload]
    w_np, wsrc = policy_weights(args.workload, args.gleet_policy)
    if w_np:
        w = {k: torch.as_tensor(v) for k, v in w_np.items()}
    else:
        w = oag.random_rlepso_weights() if args.workload == 'rlepso' else None
    cores = os.cpu_count() or 1
    fids = suite_fids(args.workload)
    runs = max(1, -(-2 * cores // len(fids)))              # >= 2 jobs per core so that every host thread stays busy
    procs = min(cores, len(fids) * runs)
    kw = dict(NP=W['NP']) if args.workload == 'lde' else {}
    gens = args.cpu_gens if args.cpu_gens is not None else (3000 if args.workload == 'deddqn' else 100)
    pool = orr.make_pool(procs)                            # created (and warmed) outside the timed region
    for _ in range(max(1, args.warmup)):
        orr.timed_sample(DIM, MAXFES, w, fids, 2, procs, suit=W['suit'], algo=args.workload, runs=runs, pool=pool, **kw)
    tot_fes, tot_s = 0, 0.0
    for k in range(args.steps):
        fes, s = orr.timed_sample(DIM, MAXFES, w, fids, gens, procs, suit=W['suit'], seed0=100 * k,
                                  algo=args.workload, runs=runs, pool=pool, **kw)
        tot_fes += fes; tot_s += s
    if pool is not None:
        pool.close(); pool.join()
    unit = 'steps' if args.workload == 'deddqn' else 'generations'
    sample = f'{len(fids)} fns x {runs} run(s) x first {gens} {unit} (dim {DIM}, pop {W["NP"]}) per step, {procs} worker processes'
    return tot_fes, tot_s, procs, 'port', sample, wsrc


def run_reference(args, rank, world):
    """The CPU arm.  kind "reference" = the unmodified reference staged under baseline/_ref (default whenever it is there); kind "port" =
    the numpy restatement under oracle/ (`--cpu-kind port`, or when baseline/_ref is absent)."""
    if rank != 0:
        return
    if 'MALLOC_MMAP_THRESHOLD_' not in os.environ:
        # numpy's [NP,NP,D] float64 temporaries (2.4 MB each) otherwise go through mmap/munmap + page faults on every generation: glibc
        # tunables, read at process start, keep them on the heap -- 1.8x faster reference rollouts, reference code untouched
        env = {**os.environ, 'MALLOC_MMAP_THRESHOLD_': '4294967296', 'MALLOC_TRIM_THRESHOLD_': '4294967296', 'MALLOC_TOP_PAD_': '268435456',
               'OMP_NUM_THREADS': '1', 'MKL_NUM_THREADS': '1', 'OPENBLAS_NUM_THREADS': '1'}
        sys.stdout.flush()
        os.execve(sys.executable, [sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env)
    W = WORKLOADS[args.workload]
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_runner as rr
    kind = args.cpu_kind
    if kind == 'auto':
        kind = 'reference' if rr.available() else 'port'
    if kind == 'reference' and not rr.available():
        raise SystemExit('baseline/_ref is not staged: run `python baseline/stage_reference.py` in the build container')
    if args.workload in ('protein', 'lde-train') and kind != 'reference':
        raise SystemExit(f'--workload {args.workload}: the CPU arm needs the staged reference (baseline/_ref); there is no oracle-port leg for it')
    if args.workload == 'lde-train':
        tot_fes, tot_s, cores, kind, sample, wsrc = run_reference_lde_train(args)
    else:
        tot_fes, tot_s, cores, kind, sample, wsrc = (run_reference_unmodified if kind == 'reference' else run_reference_port)(args)
    v = tot_fes / tot_s
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'evals/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * tot_s / max(1, args.steps),
        'higher_is_better': True, 'scaling': W.get('scaling', 'weak'), 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{W["desc"]}, dim {W.get("dim", DIM)}, pop {W["NP"]}, maxFEs {W.get("maxFEs", MAXFES)}; CPU: {sample}', 'policy': wsrc},
        'cpu_baseline': {'value': v, 'unit': 'evals/s', 'cores': cores, 'kind': kind, 'sample': sample,
                         'evals_per_s_per_core': v / max(1, cores), 'cpu': cpu_model()},
        'e2e': {'value': v, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ---------------------------------------------------------------------------------------------- our arm
def build_policy(workload, dev, NP, gleet_policy='actor'):
    import torch
    from metabox_b200 import agent as magent
    torch.manual_seed(0)
    if workload == 'gleet' and gleet_policy in ('actor', 'fused'):
        net = magent.GLEET_Actor()
        w_np, wsrc = policy_weights(workload, 'actor')
        if w_np:
            magent.load_reference_weights(net, w_np)
        else:
            wsrc = 'GLEET attention actor, random init (seed 0)'
        net = net.to(dev).eval()
        if gleet_policy == 'fused':
            return magent.GLEET_FusedActor(net), wsrc + ', fused policy kernel (mbx_gleet_policy)'
        return net, wsrc + ', torch module'
    if workload == 'gleet':
        # any torch callable state[B,NP,27] -> action[B,NP] plugs in: a small per-particle MLP with a sigmoid head
        net = torch.nn.Sequential(torch.nn.Linear(27, 64), torch.nn.Tanh(), torch.nn.Linear(64, 1), torch.nn.Sigmoid(), torch.nn.Flatten(1))
        return net.to(dev).eval(), 'stand-in per-particle MLP 27->64->1 (random init, seed 0)'
    net = {'rlepso': magent.RLEPSO_Actor, 'protein': magent.RLEPSO_Actor, 'deddqn': magent.DE_DDQN_Net}[workload]() if workload != 'lde' else magent.LDE_PolicyNet(NP=NP)
    w_np, wsrc = policy_weights(workload)
    if w_np:
        magent.load_reference_weights(net, w_np)
    return net.to(dev).eval(), wsrc


def make_sweep(workload, pack, lanes, net, seed, lane_base, NP, stripes=1, **kw):
    from metabox_b200 import sweep as msweep
    W = WORKLOADS[workload]
    maxfes, nlog = W.get('maxFEs', MAXFES), W.get('n_logpoint', 50)
    if stripes > 1:
        return msweep.StripedSweep(lambda ln, lb: make_sweep(workload, pack, ln, net, seed, lb, NP, **kw), lanes, lane_base, stripes)
    if workload in ('rlepso', 'protein'):
        return msweep.RLEPSO_Sweep(pack, lanes, net, maxfes, nlog, seed=seed, lane_base=lane_base, **kw)
    if workload == 'lde':
        return msweep.LDE_Sweep(pack, lanes, net, maxfes, nlog, seed=seed, lane_base=lane_base, NP=NP, **kw)
    if workload == 'gleet':
        return msweep.GLEET_Sweep(pack, lanes, net, maxfes, nlog, seed=seed, lane_base=lane_base, **kw)
    return msweep.DE_DDQN_Sweep(pack, lanes, net, maxfes, nlog, seed=seed, lane_base=lane_base, **kw)


def protein_data_folder():
    """The 280 protein-docking instances in the reference's text format (src/problem/protein_docking_data, 33 MB; BSD-3 data, not code):
    $MBX_PROTEIN_DATA, else the copy staged next to the reference arm (baseline/_ref, shipped with the snapshot)."""
    for d in (os.environ.get('MBX_PROTEIN_DATA'), os.path.join(ROOT, 'baseline', '_ref', 'src', 'problem', 'protein_docking_data')):
        if d and os.path.isdir(d):
            return d
    raise SystemExit('--workload protein needs the protein_docking_data folder: set MBX_PROTEIN_DATA or stage baseline/_ref '
                     '(python baseline/stage_reference.py in the build container)')


def load_problems(W):
    from metabox_b200 import problem as mp
    if W['suit'] == 'protein':
        tr, te = mp.Protein_Docking_Dataset.get_datasets('protein', difficulty='easy', data_folder=protein_data_folder())
        return tr.data + te.data
    tr, te = mp.BBOB_Dataset.get_datasets(W['suit'], W.get('dim', DIM), 5, difficulty='difficult' if W['subset'] == 'difficult-test' else 'easy')
    return te.data if W['subset'] == 'difficult-test' else tr.data + te.data


def sweep_engines(sw):
    """(lo, hi, engine) of every stripe of a sweep."""
    if hasattr(sw, 'children'):
        return [(lo, hi, c.engine) for (lo, hi), c in zip(sw.bounds, sw.children)]
    return [(0, sw.engine.B, sw.engine)]


def run_lde_train(args, rank, world, dev):
    """BASELINE config C4: the LDE *train* rollout, sharded.  One bench step = ONE learn step of `trainer.LDE_Trainer`
    (lde_agent.py:85-145 for all 30 noisy functions at once): 20 trajectories x 50 generations per function (600 lanes, sharded over the
    ranks), the REINFORCE loss + backward in torch autograd, ONE gradient all-reduce (NCCL), one Adam step.  A fixed job per step: strong
    scaling.  value = candidate evaluations charged per second; learn steps per second in config."""
    import torch
    import torch.distributed as dist
    from metabox_b200 import problem as mp, trainer as mt
    import argparse as _ap
    W = WORKLOADS['lde-train']
    NP = W['NP']
    cfg = _ap.Namespace(dim=DIM, maxFEs=MAXFES, n_logpoint=50, log_interval=MAXFES // 50)
    problems = load_problems(W)
    T = mt.LDE_Trainer(cfg, problems, NP=NP, seed=1, device=dev)
    ids = list(range(len(problems)))
    lanes = len(ids) * mt.TRAJECTORY_NUM
    evals_per_step = lanes * (NP + mt.TRAJECTORY_LENGTH * NP)              # every trajectory: init_population + 50 generations

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(int(os.environ.get('LOCAL_RANK', 0)))
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        T.train_step(ids)
    barrier()
    clocks.mark()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        info = T.train_step(ids)
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    # e2e: the instance tables come from pinned host memory every step, the loss goes back to the host (train_step ends with .item())
    h_hdr = torch.from_numpy(np.frombuffer(T.pack.h_headers.tobytes(), dtype=np.uint8).copy()).pin_memory()
    h_cst = torch.from_numpy(T.pack.h_consts).pin_memory()

    def e2e_step():
        T.pack.headers.copy_(h_hdr, non_blocking=True)
        T.pack.consts.copy_(h_cst, non_blocking=True)
        return T.train_step(ids)['loss']

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e_ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    # kernel share: one instrumented collection on rank 0's shard (CUDA events around every fused-step launch)
    lo, hi = (0, len(ids))
    if world > 1:
        from metabox_b200 import sweep as msweep
        lo, hi = msweep.shard_range(len(ids), rank, world)
    k_ms, k_n, c_ms = 0.0, 0, 0.0
    if rank == 0 and hi > lo:
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        tr = T.collect(ids[lo:hi], lane_base=lo * mt.TRAJECTORY_NUM, time_kernels=True)
        c1.record(); torch.cuda.synchronize()
        k_ms, k_n, c_ms = sum(a.elapsed_time(b) for a, b in tr['kernel_events']), len(tr['kernel_events']), c0.elapsed_time(c1)
    tt = torch.tensor([ms_total, e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, e_ms = float(tt[0]), float(tt[1])
    barrier()
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_loc = (hi - lo) * mt.TRAJECTORY_NUM
        k_fes = n_loc * mt.TRAJECTORY_LENGTH * NP
        achieved = k_fes * W['bytes_per_eval'] / (k_ms / 1000.0) / 1e9 if k_ms > 0 else None
        n_par = sum(p.numel() for p in T.net.parameters())
        out = {
            'metric': METRIC, 'value': evals_per_step * args.steps / (ms_total / 1000.0), 'unit': 'evals/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'{W["desc"]}; dim {DIM}, pop {NP}; {lanes} lanes sharded over {world} GPU(s) (strong scaling: a fixed job per '
                                   f'learn step); one step = collect trajectories + REINFORCE backward + ONE gradient all-reduce + Adam step',
                       'learn_steps_per_s': args.steps / (ms_total / 1000.0), 'lanes_per_gpu': n_loc,
                       'policy': f'LDE LSTM policy, NP {NP}, random init (the shipped policy only fits NP 50), trained in place ({n_par} parameters)',
                       'collective': f'one all-reduce of {n_par + 1} float32 per learn step (gradients + sample count)',
                       'launch': 'eager: 50 x (LSTM policy step in torch, fused LDE step kernel) per learn step, then torch autograd',
                       'l2': 'per-GPU state a few MB (L2 resident; the step is launch-latency bound at 600 lanes)'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': (achieved / peak) if achieved else None,
                         'traffic': None, 'kernel': W['kernel'], 'avg_launch_ms': k_ms / max(1, k_n), 'launches_timed': k_n,
                         'kernel_share_of_step': (k_ms / (ms_total / args.steps)) if ms_total else None,
                         'collect_ms': c_ms, 'peak_source': peak_src,
                         'note': f'algorithmic bytes = {W["bytes_per_eval"]} B/eval x evals of the timed step launches (rank 0 shard); ' + W['bound_note']},
            'e2e': {'value': evals_per_step * args.steps / (e_ms / 1000.0), 'unit': 'evals/s',
                    'h2d_bytes_per_step': int(h_hdr.numel() + h_cst.numel() * 8), 'd2h_bytes_per_step': 4},
            'gpu_launches': int(args.steps * (mt.TRAJECTORY_LENGTH + 1)), 'clocks': clk,
            'last_info': {k: info[k] for k in ('loss', 'return', 'learn_steps', 'problems')},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                                    '--workload', 'lde-train'], capture_output=True, text=True, timeout=900,
                                   env={**os.environ, 'RANK': '0', 'WORLD_SIZE': '1', 'CUDA_VISIBLE_DEVICES': ''})
                out['cpu_baseline'] = json.loads(r.stdout.strip().splitlines()[-1])['cpu_baseline']
            except Exception as ex:
                out['cpu_baseline'] = {'value': None, 'unit': 'evals/s', 'cores': 0, 'kind': 'reference', 'sample': f'failed: {ex}'}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    from metabox_b200 import problem as mp, sweep as msweep, _lib
    _lib.lib()
    W = WORKLOADS[args.workload]
    NP = W['NP']

    if args.workload == 'lde-train':
        return run_lde_train(args, rank, world, dev)
    dim, maxfes = W.get('dim', DIM), W.get('maxFEs', MAXFES)
    strong = W.get('scaling', 'weak') == 'strong'
    problems = load_problems(W)
    pack = mp.InstancePack(problems, dev)                      # also computes every optimum on the GPU
    if args.torch_policy and args.workload == 'gleet' and args.gleet_policy == 'fused':
        args.gleet_policy = 'actor'
    net, wsrc = build_policy(args.workload, dev, NP, args.gleet_policy)
    if not args.torch_policy and args.workload == 'deddqn':
        from metabox_b200 import agent as magent
        net, wsrc = magent.DE_DDQN_FusedNet(net), wsrc + ', fused Q-network kernel (mbx_deddqn_policy)'
    if not args.torch_policy and args.workload in ('rlepso', 'protein'):
        from metabox_b200 import agent as magent
        net, wsrc = magent.RLEPSO_FusedActor(net), wsrc + ', fused actor kernel (mbx_rlepso_policy)'
    inst, run = msweep.lane_table(len(problems), args.runs)
    if args.lane_order == 'run-major':                        # every contiguous stripe / rank sees the same mix of functions
        inst = np.ascontiguousarray(inst.reshape(len(problems), args.runs).T.reshape(-1))
    elif args.lane_order == 'stripe-balanced':                # every stripe gets ~runs/stripes runs of EVERY function, instance-major
        parts = [np.array_split(np.arange(args.runs), args.stripes) for _ in problems]          # inside (constants stay L1/L2-hot)
        inst = np.concatenate([np.full(len(parts[k][s_]), k, dtype=np.int32) for s_ in range(args.stripes) for k in range(len(problems))])
    n_global = len(inst)
    if strong:                                                # a fixed job: the ranks shard its lanes (contiguous blocks, sizes differ by <= 1)
        lo, hi = msweep.shard_range(n_global, rank, world)
        inst, lane_base = inst[lo:hi], lo
        n_max = -(-n_global // world)
    else:                                                     # weak scaling: every rank gets its own n_fns x runs lanes
        lane_base = rank * n_global
        n_max = n_global
    n_local = len(inst)
    seed = 20261017
    init_fes = n_local * NP                                   # evaluations charged by init_population (not by step launches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_sweep(sw):
        res = sw.run()
        rows = msweep.gather_results(res, world, n_max=n_max)     # the single collective of the sweep
        return res, rows

    # ---- value: inputs resident in HBM; policy forward + fused step replayed from a CUDA graph ------------------
    gkw = {} if args.steps_per_graph is None else {'steps_per_graph': args.steps_per_graph}
    if args.poll is not None:
        gkw['poll'] = args.poll
    if args.workload == 'deddqn' and not args.torch_policy and args.launches_per_poll is not None:
        gkw['launches_per_poll'] = args.launches_per_poll
    sw = make_sweep(args.workload, pack, inst, net, seed, lane_base, NP, stripes=args.stripes, graph=not args.no_graph, **gkw)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        one_sweep(sw)
    barrier()
    clocks.mark()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    fes_total, launches = 0, 0
    for _ in range(args.steps):
        res, rows = one_sweep(sw)
        fes_total += int(res['fes'].sum().item())
        launches += sw.launches
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    tt = torch.tensor([ms_total, float(fes_total)], device=dev, dtype=torch.float64)
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, fes_all = float(mx[0]), float(sm[1])
    else:
        fes_all = float(fes_total)
    value = fes_all / (ms_total / 1000.0)

    # ---- e2e: host buffers in (packed instance table + lane table from pinned memory), host rows out -------------
    h_hdr = torch.from_numpy(np.frombuffer(pack.h_headers.tobytes(), dtype=np.uint8).copy()).pin_memory()
    h_cst = torch.from_numpy(pack.h_consts).pin_memory()
    h_lane = torch.from_numpy(inst).pin_memory()
    n_rows = n_global if strong else n_local * world
    h_rows = torch.empty(n_rows, 53, dtype=torch.float32).pin_memory()
    h2d = h_hdr.numel() + h_cst.numel() * 8 + h_lane.numel() * 4
    d2h = h_rows.numel() * 4

    def e2e_step():
        pack.headers.copy_(h_hdr, non_blocking=True)           # H2D of this step's inputs
        pack.consts.copy_(h_cst, non_blocking=True)
        for lo, hi, eng in sweep_engines(sw):                   # lane table of every stripe
            eng.lane_inst.copy_(h_lane[lo:hi], non_blocking=True)
        res = sw.run()
        rows = msweep.gather_results(res, world, n_max=n_max)
        h_rows.copy_(rows, non_blocking=True)                  # D2H of the step's result
        torch.cuda.synchronize()
        return float(h_rows[:, 51].sum())

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e_fes = 0.0
    for _ in range(args.steps):
        e_fes += e2e_step()
    e1.record()
    barrier()
    e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e_ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e_ms = float(t[0])
    e2e_value = e_fes / (e_ms / 1000.0)        # h_rows holds all ranks' lanes
    clk = clocks.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel: one instrumented sweep of the same workload, CUDA events around every step
    #      launch on the launching stream (a graph replay cannot be timed per kernel) ---------------------------------
    k_ms = k_n = k_fes = 0
    k_sweep_ms = 0.0
    if rank == 0:
        swi = make_sweep(args.workload, pack, inst, net, seed, lane_base, NP, time_kernels=True)
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        res = swi.run()
        i1.record()
        k_ms, k_n = swi.kernel_ms()
        k_sweep_ms = i0.elapsed_time(i1)
        k_fes = int(res['fes'].sum().item()) - init_fes
        del swi
    barrier()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = (k_fes * W['bytes_per_eval']) / (k_ms / 1000.0) / 1e9 if k_ms > 0 else None
        traffic, traffic_note, tr = None, 'no ncu --set full capture recorded for this kernel', None
        try:                                                    # DRAM bytes per launch from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get(W['kernel'])
            if tr:
                traffic = tr['dram_bytes_per_launch'] * n_local / tr['lanes_per_launch']
                traffic_note = (f"bytes per launch of {n_local} lanes, scaled from {tr['dram_bytes_per_launch']:.4g} B at "
                                f"{tr['lanes_per_launch']} lanes ({tr['launch']}); {tr['source']}; algorithmic bytes of a full launch = "
                                f"{n_local * NP * W['bytes_per_eval']:.4g}")
        except Exception:
            pass
        # second roofline for kernels that are bound by instruction ISSUE, not bytes: warp instructions per evaluation from the committed ncu
        # capture x the measured evaluation rate of the step launches, against 4 issue slots per SM-cycle at the sampled SM clock
        issue = None
        try:
            if tr and tr.get('warp_instructions_per_launch') and k_ms > 0:
                wipe = tr['warp_instructions_per_launch'] / tr['evals_per_launch']
                props = torch.cuda.get_device_properties(dev)
                mhz = (clk or {}).get('sm_mhz') or 1965.0
                peak_issue = 4.0 * props.multi_processor_count * mhz * 1e6
                ach = (k_fes / (k_ms / 1000.0)) * wipe
                issue = {'warp_instructions_per_eval': wipe, 'achieved_warp_instr_per_s': ach, 'peak_warp_instr_per_s': peak_issue,
                         'frac': ach / peak_issue, 'source': tr.get('issue_source')}
        except Exception:
            issue = None
        state_mb = {'rlepso': n_local * (3 * NP * dim + 3 * NP) * 4, 'protein': n_local * (3 * NP * dim + 3 * NP) * 4,
                    'lde': n_local * (NP * dim + NP) * 4, 'gleet': n_local * (3 * NP * dim + 3 * NP + 9 * NP) * 4,
                    'deddqn': n_local * ((NP * dim + NP) * 4 + 8800)}[args.workload] / 1e6
        spg = args.steps_per_graph or (10 if args.workload == 'deddqn' else 1)
        if args.no_graph:
            launch = 'eager'
        elif args.workload == 'deddqn' and not args.torch_policy:
            launch = (f'persistent kernel (mbx_deddqn_run): {args.poll or 250} function evaluations per launch, {args.launches_per_poll or 16} launches between host polls, '
                      f'Q-network inside the kernel, lane state resident in shared memory; {args.stripes} lane stripe(s) on separate streams')
        else:
            launch = f'policy forward + fused step replayed from one CUDA graph per {spg} step(s); {args.stripes} lane stripe(s) on separate streams'
        lanes_txt = (f'{len(problems)} instances x {args.runs} runs = {n_global} lanes sharded over {world} GPU(s) (strong scaling: a fixed job)'
                     if strong else f'x {args.runs} runs/GPU')
        out = {
            'metric': METRIC, 'value': value, 'unit': 'evals/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': W.get('scaling', 'weak'), 'vs_baseline': None, 'dtype': 'f64',
            'data': 'real instances (reference data set), Philox populations' if args.workload == 'protein' else 'synthetic',
            'config': {'workload': f'{W["desc"]} {lanes_txt}, dim {dim}, pop {NP}, maxFEs {maxfes}, '
                                   'one step = one complete sweep (init + all generations + all-gather)',
                       'lanes_per_gpu': n_local, 'lane_order': args.lane_order, 'state': 'float32 HBM / float64 registers', 'policy': wsrc,
                       'launch': launch,
                       'l2': f'per-GPU state {state_mb:.0f} MB vs 126 MB L2' + (' (inputs larger than L2)' if state_mb > 126 else ' (L2 resident; the kernel is not DRAM bound)'),
                       'parallelism': f'lanes sharded over {world} GPU(s), one NCCL all-gather per sweep',
                       'e2e': 'the plug-in call with HOST buffers: the packed instance table + lane table go host -> device from pinned memory and the '
                              'result rows (cost curve, fes, return per lane) device -> host inside the timed region, every step; populations '
                              'are born on the device (Philox), so there are no candidate buffers to upload'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': (achieved / peak) if achieved else None, 'traffic': traffic, 'traffic_note': traffic_note,
                         'kernel': W['kernel'], 'avg_launch_ms': k_ms / max(1, k_n), 'launches_timed': k_n,
                         'kernel_share_of_step': (k_ms / k_sweep_ms) if k_sweep_ms else None, 'peak_source': peak_src, 'issue': issue,
                         'timed': 'separate instrumented (eager, unstriped) sweep of the same workload right after the timed region; '
                                  'kernel_share_of_step = step-kernel time / duration of that sweep',
                         'note': f'algorithmic bytes = {W["bytes_per_eval"]} B/eval x evals charged by the step launches; see DESIGN.md section 4 '
                                 'for what actually bounds the kernel: ' + W['bound_note']},
            'e2e': {'value': e2e_value, 'unit': 'evals/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'gpu_launches': int(launches), 'clocks': clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            # the CPU arm, timed in fresh processes (no CUDA context, so their fork pools are safe): the unmodified reference when it is
            # staged (kind "reference"), and the oracle port as a second, labelled figure
            def cpu_leg(kind, steps):
                cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', str(steps), '--warmup', '1', '--cpu-kind', kind,
                       '--workload', args.workload, '--gleet-policy', 'stand-in' if args.gleet_policy == 'stand-in' else 'actor']
                if args.cpu_gens is not None:
                    cmd += ['--cpu-gens', str(args.cpu_gens)]
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                                   env={**os.environ, 'RANK': '0', 'WORLD_SIZE': '1', 'CUDA_VISIBLE_DEVICES': ''})
                return json.loads(r.stdout.strip().splitlines()[-1])['cpu_baseline']
            try:
                out['cpu_baseline'] = cpu_leg('auto', 1)
            except Exception as ex:                              # never lose the GPU line over the baseline
                out['cpu_baseline'] = {'value': None, 'unit': 'evals/s', 'cores': 0, 'kind': 'port', 'sample': f'failed: {ex}'}
            if out['cpu_baseline'].get('kind') == 'reference':
                try:
                    out['cpu_baseline_port'] = cpu_leg('port', 2)
                except Exception as ex:
                    out['cpu_baseline_port'] = {'value': None, 'sample': f'failed: {ex}'}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
