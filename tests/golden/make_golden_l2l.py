#!/usr/bin/env python
"""Golden episodes of the UNMODIFIED reference L2L agent + optimizer (src/agent/l2l_agent.py `rollout_episode`, src/optimizer/l2l_optimizer.py),
build container only: python tests/golden/make_golden_l2l.py -> tests/golden/l2l.npz.  The shipped agent (`agent_model/test/bbob_easy/
L2L_Agent.pkl`, dim 10) rolls out on F32Problem proxies; the episode is deterministic (no random draws)."""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402
import torch  # noqa: E402

import make_golden as mg  # noqa: E402

FIDS = [1, 5, 8, 15, 21, 104]


def main():
    import agent  # noqa: F401
    from environment import PBO_Env
    from optimizer import L2L_Optimizer
    with open(os.path.join(_refenv.REF, 'src', 'agent_model', 'test', 'bbob_easy', 'L2L_Agent.pkl'), 'rb') as f:
        ag = pickle.load(f)
    out = dict(fids=np.array(FIDS))
    for k, v in ag.net.state_dict().items():
        out['net/' + k] = v.numpy()
    for fid in FIDS:
        prob = _refenv.F32Problem(mg.bbob_instance(fid, 10))
        ys = []
        ev0 = prob.eval
        prob.eval = lambda x, ev0=ev0, ys=ys: (lambda f: (ys.append(float(f)), np.float64(f))[1])(ev0(x))      # record every cost the episode sees
        np.random.seed(fid)                                                                         # (noisy functions draw from np.random)
        res = ag.rollout_episode(PBO_Env(prob, L2L_Optimizer(mg.ref_config(10))))
        out[f'f{fid}/y'] = np.array(ys); out[f'f{fid}/cost'] = np.array([float(c) for c in res['cost']])
        out[f'f{fid}/fes'] = np.int64(res['fes']); out[f'f{fid}/ret'] = np.float64(res['return'])
        print(fid, 'fes', res['fes'], 'best', min(ys), 'return', float(res['return']), 'len(cost)', len(res['cost']))
    torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, 'l2l.npz'), **out)


if __name__ == '__main__':
    main()
