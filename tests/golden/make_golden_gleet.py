#!/usr/bin/env python
"""Golden trajectories of the UNMODIFIED reference `GLEET_Optimizer` (src/optimizer/gleet_optimizer.py), run in the
build container only (needs /root/reference):  python tests/golden/make_golden_gleet.py  ->  tests/golden/gleet.npz

Teacher forcing like make_golden.py's RLEPSO section: the reference optimizer runs on an `F32Problem` proxy and its
state (positions, velocities, pbest, costs, carried features) is rounded to float32-representable values after every
step, so the reference itself follows the engine's storage contract without touching its code."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402
import torch  # noqa: E402

import make_golden as mg  # noqa: E402  (bbob_instance / protein_instance / ref_config / f32r)

CASES = [('f1_d10', 1, 10, 6), ('f3_d10', 3, 10, 5), ('f8_d10', 8, 10, 4), ('f15_d30', 15, 30, 4), ('f21_d10', 21, 10, 4),
         ('f104_d10', 104, 10, 4), ('f124_d10', 124, 10, 4), ('prot_1R0R_1', '1R0R_1', 12, 4)]


def get(opt):
    P = opt.particles
    return dict(x=P['current_position'], v=P['velocity'], pbest_x=P['pbest_position'], c_cost=P['c_cost'], pbest=P['pbest'],
                gbest_x=P['gbest_position'], gbest_val=np.float64(P['gbest_val']), gbest_idx=np.int64(P['gbest_index']),
                pni=np.asarray(opt.per_no_improve).astype(np.int32), no_improve=np.int64(opt.no_improve), w=np.float64(opt.w),
                max_cost=np.float64(opt.max_cost), pbest_feature=opt.pbest_feature, gbest_feature=opt.gbest_feature,
                fes=np.int64(opt.fes), log_index=np.int64(opt.log_index), curve=np.array(opt.cost, dtype=np.float64))


def round_state(opt):
    P = opt.particles
    for k in ('current_position', 'velocity', 'pbest_position', 'c_cost', 'pbest', 'gbest_position'):
        P[k] = mg.f32r(P[k])
    P['gbest_val'] = np.float64(np.float32(P['gbest_val']))
    opt.max_cost = np.float64(np.float32(opt.max_cost))
    opt.pbest_feature = mg.f32r(opt.pbest_feature)
    opt.gbest_feature = mg.f32r(opt.gbest_feature)


def store(out, pre, st):
    for k, v in st.items():
        f32 = getattr(v, 'ndim', 0) and v.dtype == np.float64 and k != 'curve'
        out[pre + k] = v.astype(np.float32) if f32 else v


def main():
    from optimizer import GLEET_Optimizer
    out = dict(meta=str(dict(numpy=np.__version__, torch=torch.__version__)), cases=np.array([c[0] for c in CASES]))
    ga = np.random.RandomState(777)
    for key, fid, dim, n_steps in CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, 1000, 5)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim)
        prob = _refenv.F32Problem(inner)
        opt = GLEET_Optimizer(cfg)
        prob.reset()
        seed0 = 5000 + len(key) * 11
        np.random.seed(seed0)
        state = opt.init_population(prob)
        round_state(opt)
        out[f'{key}/fid'] = np.array(str(fid)); out[f'{key}/dim'] = np.int64(dim)
        out[f'{key}/maxFEs'] = np.int64(cfg.maxFEs); out[f'{key}/n_logpoint'] = np.int64(cfg.n_logpoint)
        out[f'{key}/seed0'] = np.int64(seed0); out[f'{key}/n_steps'] = np.int64(n_steps)
        store(out, f'{key}/s0/', get(opt)); out[f'{key}/s0/state'] = np.asarray(state, dtype=np.float64)
        for t in range(1, n_steps + 1):
            action = ga.rand(100).astype(np.float32)
            if t == 2:
                action[:10] = 0.0; action[10:20] = 1.0                 # the clamp ends of the action range
            seed = seed0 + t
            np.random.seed(seed)
            state, reward, done = opt.update(action, prob)
            round_state(opt)
            out[f'{key}/s{t}/action'] = action; out[f'{key}/s{t}/seed'] = np.int64(seed)
            out[f'{key}/s{t}/state'] = np.asarray(state, dtype=np.float64)
            out[f'{key}/s{t}/reward'] = np.float64(reward); out[f'{key}/s{t}/done'] = np.bool_(done)
            store(out, f'{key}/s{t}/', get(opt))
    np.savez_compressed(os.path.join(HERE, 'gleet.npz'), **out)
    print('gleet done', os.path.getsize(os.path.join(HERE, 'gleet.npz')))


if __name__ == '__main__':
    main()
