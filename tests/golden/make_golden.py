#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py [section ...]
Sections: bbob protein rlepso rs lde deddqn agents.   Fixtures are small .npz files; each records the
numpy / torch versions used.  The tests never import the reference -- they read these files.

What each fixture pins
  bbob_d{10,30}.npz     reference `problem.eval` (noise free) for all 54 function ids on float32-
                        representable points (uniform, near-optimum, out-of-bounds), instance
                        constants for D=10, and noisy `eval` under fixed np.random seeds.
  protein.npz           6 shipped complexes (raw per-atom data) + reference energies.
  rlepso.npz            teacher-forced `RLEPSO_Optimizer.init_population/update` trajectories run by the
                        reference optimizer class on `F32Problem` proxies (storage contract, see
                        _refenv.py); random draws are reproduced from the recorded np.random seeds.
  rs.npz                `Random_search.run_episode` results.
"""
import argparse
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402
import torch  # noqa: E402

from problem import bbob as rbbob  # noqa: E402
from problem import protein_docking as rprot  # noqa: E402

META = dict(numpy=np.__version__, torch=torch.__version__)


def f32r(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


# ------------------------------------------------------------------------------------------ bbob
def points_for(inst, dim, g):
    uni = g.uniform(-5, 5, size=(16, dim))
    near = inst.opt[None, :] + 1e-3 * g.normal(size=(8, dim))
    mid = inst.opt[None, :] + 0.3 * g.normal(size=(4, dim))
    oob = g.uniform(-6.5, 6.5, size=(4, dim))
    spec = np.stack([inst.opt, np.zeros(dim), np.full(dim, 5.0), np.full(dim, -5.0)])
    return f32r(np.concatenate([uni, near, mid, oob, spec], 0))


def gen_bbob():
    for dim in (10, 20, 30):                               # 20: the third compiled dimension (csrc/mbx_launch.h MBX_FOR_BBOB_DIMS)
        out = dict(meta=str(META))
        for suit in ('bbob', 'bbob-noisy'):
            tr, te = rbbob.BBOB_Dataset.get_datasets(suit=suit, dim=dim, upperbound=5, difficulty='easy')
            insts = tr.data + te.data
            ids = [int(type(p).__name__[1:]) for p in insts]
            g = np.random.RandomState(1234 + dim)
            for fid, p in sorted(zip(ids, insts)):
                X = points_for(p, dim, g)
                p.reset()
                if suit == 'bbob':
                    F = p.eval(X)
                else:
                    F = rbbob.BBOB_Basic_Problem.eval(p, X)            # noise-free ftrue
                    np.random.seed(777 + fid)
                    out[f'fn_{fid}'] = p.eval(X)                       # with noise, seed 777+fid
                    np.random.seed(778 + fid)
                    out[f'fn1_{fid}'] = np.array([p.eval(X[k]) for k in range(4)])   # 1-D dispatch
                out[f'x_{fid}'] = X.astype(np.float32)
                out[f'f_{fid}'] = F
                out[f'opt_{fid}'] = np.float64(p.optimum)
                out[f'name_{fid}'] = str(p)
                if dim == 10:
                    out[f'shift_{fid}'] = p.shift
                    out[f'R_{fid}'] = p.rotate
                    out[f'bias_{fid}'] = np.float64(p.bias)
                    for nm, key in (('Q_rotate', 'M2'), ('linearTF', 'M2'), ('y', 'y'), ('C', 'C'), ('w', 'w')):
                        if hasattr(p, nm):
                            out[f'{key}_{fid}'] = np.asarray(getattr(p, nm))
            trn, ten = rbbob.BBOB_Dataset.get_datasets(suit=suit, dim=dim, upperbound=5, difficulty='difficult')
            out[f'split_easy_{suit}'] = np.array([str(p) for p in te.data])
            out[f'split_difficult_{suit}'] = np.array([str(p) for p in ten.data])
        np.savez_compressed(os.path.join(HERE, f'bbob_d{dim}.npz'), **out)
        print('bbob', dim, 'done')


# ------------------------------------------------------------------------------------------ protein
PROTEIN_IDS = ['1AVX_1', '1R0R_1', '2JEL_7', '1GRN_3', '1ATN_10', '2C0L_5']


def gen_protein():
    out = dict(meta=str(META), ids=np.array(PROTEIN_IDS))
    folder = os.path.join(_refenv.REF, 'src', 'problem', 'protein_docking_data')
    g = np.random.RandomState(99)
    for pid in PROTEIN_IDS:
        d = os.path.join(folder, pid)
        coor = np.loadtxt(d + '/coor_init'); q = np.loadtxt(d + '/q'); e = np.loadtxt(d + '/e')
        r = np.loadtxt(d + '/r'); basis = np.loadtxt(d + '/basis'); eig = np.loadtxt(d + '/eigval')
        qq = np.matmul(np.tile(q, (1, 1)).T, np.tile(q, (1, 1)))
        ee = np.sqrt(np.matmul(np.tile(e, (1, 1)).T, np.tile(e, (1, 1))))
        rr = np.tile(r, (len(r), 1)); rr = (rr + rr.T) / 2
        p = rprot.Protein_Docking(coor, qq, ee, rr, basis, eig, pid)
        p.reset()
        X = f32r(np.concatenate([g.uniform(-1.5, 1.5, size=(24, 12)), 0.05 * g.normal(size=(7, 12)),
                                 np.zeros((1, 12))], 0))
        out[f'x_{pid}'] = X.astype(np.float32)
        out[f'energy_{pid}'] = p.eval(X)
        for k, v in dict(coor_init=coor, q=q, e=e, r=r, basis=basis, eigval=eig).items():
            out[f'{k}_{pid}'] = v
    # dataset split order (protein_docking.py:138-160)
    tr, te = rprot.Protein_Docking_Dataset.get_datasets('protein', difficulty='easy')
    out['split_easy_train'] = np.array([str(p) for p in tr.data])
    out['split_easy_test'] = np.array([str(p) for p in te.data])
    np.savez_compressed(os.path.join(HERE, 'protein.npz'), **out)
    print('protein done')


# ------------------------------------------------------------------------------------------ helpers
def ref_config(dim, maxFEs=None, n_logpoint=50):
    maxFEs = maxFEs or 2000 * dim
    return argparse.Namespace(dim=dim, maxFEs=maxFEs, n_logpoint=n_logpoint, log_interval=maxFEs // n_logpoint,
                              upperbound=5, device='cpu')


def bbob_instance(fid, dim):
    suit = 'bbob' if fid < 100 else 'bbob-noisy'
    tr, te = rbbob.BBOB_Dataset.get_datasets(suit=suit, dim=dim, upperbound=5, difficulty='easy')
    for p in tr.data + te.data:
        if type(p).__name__ == f'F{fid}':
            return p
    raise KeyError(fid)


def protein_instance(pid):
    folder = os.path.join(_refenv.REF, 'src', 'problem', 'protein_docking_data')
    d = os.path.join(folder, pid)
    coor = np.loadtxt(d + '/coor_init'); q = np.loadtxt(d + '/q'); e = np.loadtxt(d + '/e')
    r = np.loadtxt(d + '/r'); basis = np.loadtxt(d + '/basis'); eig = np.loadtxt(d + '/eigval')
    qq = np.matmul(np.tile(q, (1, 1)).T, np.tile(q, (1, 1)))
    ee = np.sqrt(np.matmul(np.tile(e, (1, 1)).T, np.tile(e, (1, 1))))
    rr = np.tile(r, (len(r), 1)); rr = (rr + rr.T) / 2
    return rprot.Protein_Docking(coor, qq, ee, rr, basis, eig, pid)


# ------------------------------------------------------------------------------------------ rlepso
RLEPSO_CASES = [  # (key, fid or protein id, dim, n_steps)
    ('f1_d10', 1, 10, 6), ('f3_d10', 3, 10, 6), ('f15_d10', 15, 10, 6), ('f21_d10', 21, 10, 4),
    ('f20_d10', 20, 10, 4), ('f15_d30', 15, 30, 4), ('f103_d10', 103, 10, 5), ('f114_d10', 114, 10, 4),
    ('prot_1R0R_1', '1R0R_1', 12, 4),
]


def _rlepso_get(opt):
    P = opt._RLEPSO_Optimizer__particles
    return dict(x=P['current_position'], v=P['velocity'], pbest_x=P['pbest_position'], c_cost=P['c_cost'],
                pbest=P['pbest'], gbest_x=P['gbest_position'], gbest_val=np.float64(P['gbest_val']),
                gbest_idx=np.int64(P['gbest_index']), pni=opt._RLEPSO_Optimizer__per_no_improve.copy(),
                fes=np.int64(opt.fes), log_index=np.int64(opt.log_index), curve=np.array(opt.cost, dtype=np.float64))


def _rlepso_round(opt):
    """Teacher forcing: round the reference optimizer's float64 state to float32-representable."""
    P = opt._RLEPSO_Optimizer__particles
    for k in ('current_position', 'velocity', 'pbest_position', 'c_cost', 'pbest', 'gbest_position'):
        P[k] = f32r(P[k])
    P['gbest_val'] = np.float64(np.float32(P['gbest_val']))


def gen_rlepso():
    from optimizer import RLEPSO_Optimizer
    out = dict(meta=str(META), cases=np.array([c[0] for c in RLEPSO_CASES]))
    ga = np.random.RandomState(4242)
    for key, fid, dim, n_steps in RLEPSO_CASES:
        if isinstance(fid, str):
            inner = protein_instance(fid); cfg = ref_config(12, 1000, 5)
        else:
            inner = bbob_instance(fid, dim); cfg = ref_config(dim)
        prob = _refenv.F32Problem(inner)
        opt = RLEPSO_Optimizer(cfg)
        prob.reset()
        seed0 = 1000 + len(key) * 7
        np.random.seed(seed0)
        opt.init_population(prob)
        _rlepso_round(opt)
        st = _rlepso_get(opt)
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim)
        out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs); out[f'{key}/n_logpoint'] = np.int64(cfg.n_logpoint)
        out[f'{key}/seed0'] = np.int64(seed0); out[f'{key}/n_steps'] = np.int64(n_steps)
        for k, v in st.items():
            out[f'{key}/s0/{k}'] = v.astype(np.float32) if getattr(v, 'ndim', 0) and v.dtype == np.float64 and k != 'curve' else v
        for t in range(1, n_steps + 1):
            action = ga.rand(35).astype(np.float32)
            if t == 3:      # provoke the stochastic re-initialisation branch (rlepso_optimizer.py:238)
                opt._RLEPSO_Optimizer__per_no_improve[:] = 40.0 + np.arange(100) % 7
                action[::5] = np.maximum(action[::5], 0.6)
            pni_in = opt._RLEPSO_Optimizer__per_no_improve.copy()
            seed = seed0 + t
            np.random.seed(seed)
            state, reward, done = opt.update(action, prob)
            _rlepso_round(opt)
            st = _rlepso_get(opt)
            out[f'{key}/s{t}/action'] = action; out[f'{key}/s{t}/seed'] = np.int64(seed)
            out[f'{key}/s{t}/pni_in'] = pni_in.astype(np.int32)
            out[f'{key}/s{t}/feature'] = np.asarray(state, dtype=np.float64)
            out[f'{key}/s{t}/reward'] = np.int64(reward); out[f'{key}/s{t}/done'] = np.bool_(done)
            for k, v in st.items():
                out[f'{key}/s{t}/{k}'] = v.astype(np.float32) if getattr(v, 'ndim', 0) and v.dtype == np.float64 and k != 'curve' else v
    np.savez_compressed(os.path.join(HERE, 'rlepso.npz'), **out)
    print('rlepso done')


# ------------------------------------------------------------------------------------------ random search
def gen_rs():
    from optimizer import Random_search
    out = dict(meta=str(META))
    for fid, dim in ((1, 10), (15, 10), (8, 30), (102, 10)):
        inner = bbob_instance(fid, dim)
        cfg = ref_config(dim)
        for contract in ('f64', 'f32'):
            prob = inner if contract == 'f64' else _refenv.F32Problem(inner)
            np.random.seed(5)
            res = Random_search(cfg).run_episode(prob)
            out[f'{contract}/{fid}_{dim}/cost'] = np.array(res['cost'], dtype=np.float64)
            out[f'{contract}/{fid}_{dim}/fes'] = np.int64(res['fes'])
    np.savez_compressed(os.path.join(HERE, 'rs.npz'), **out)
    print('rs done')


SECTIONS = dict(bbob=gen_bbob, protein=gen_protein, rlepso=gen_rlepso, rs=gen_rs)

if __name__ == '__main__':
    try:
        import make_golden_de  # noqa: F401  (adds lde / deddqn / agents sections)
        SECTIONS.update(make_golden_de.SECTIONS)
    except ImportError:
        pass
    which = sys.argv[1:] or list(SECTIONS)
    for s in which:
        SECTIONS[s]()
