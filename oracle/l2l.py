"""Oracle: L2L backbone + agent rollout (a learned point-wise optimizer: an LSTM proposes ONE point per step, 100 steps).  TEST INFRASTRUCTURE ONLY.

numpy / torch-CPU restatement of `src/optimizer/l2l_optimizer.py:10-53` (`np_scale`, `update`) and `src/agent/l2l_agent.py:97-134`
(`rollout_episode`: LSTM(dim + 2 -> 32, proj dim) in float64, input = [previous raw output | previous cost | 1], zero initial input / state)
under the engine's storage contract (the proposed point is rounded to float32 before evaluation, the cost `eval(x) - optimum` to float32).
The rollout has no random draws at all, so whole episodes are comparable.  Pinned by `tests/golden/l2l.npz`.
"""
from __future__ import annotations

import numpy as np
import torch

T_MAX = 100                        # l2l_optimizer.py:50, l2l_agent.py:99


def np_scale(x, lb, ub):
    """l2l_optimizer.py:10-13"""
    x = 1 / (1 + np.exp(-x))
    return lb + (ub - lb) * x


def step_cost(ev, optimum, lb, ub, raw):
    """update (:26-40): raw LSTM output [D] f64 -> cost of the scaled, float32-rounded point (float32-representable float64)."""
    x32 = np_scale(np.asarray(raw, dtype=np.float64), lb, ub).astype(np.float32).astype(np.float64)
    f = ev(x32[None, :], None)
    y = f if optimum is None else f - optimum
    return float(np.asarray(y, dtype=np.float64).astype(np.float32)[0])


@torch.no_grad()
def lstm_step(w, inp, h, c):
    """One step of nn.LSTM(dim + 2, 32, proj_size=dim), float64; gate order i, f, g, o; h' = W_hr (o * tanh(c'))."""
    gates = w['weight_ih_l0'] @ inp + w['bias_ih_l0'] + w['weight_hh_l0'] @ h + w['bias_hh_l0']
    i, f, g, o = gates.chunk(4)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = w['weight_hr_l0'] @ (torch.sigmoid(o) * torch.tanh(c2))
    return h2, c2


def rollout(w, dim, ev, optimum, lb, ub):
    """rollout_episode (:97-134) -> dict(raw [n,D], y [n], cost (every 2nd entry of the best-so-far list), fes, ret = sum(y) / y[0])."""
    inp = torch.zeros(dim + 2, dtype=torch.float64)
    h, c = torch.zeros(dim, dtype=torch.float64), torch.zeros(w['weight_hr_l0'].shape[1], dtype=torch.float64)
    raws, ys, best_list, best = [], [], [], None
    for t in range(T_MAX):
        h, c = lstm_step(w, inp, h, c)
        raw = h.numpy().copy()
        y = step_cost(ev, optimum, lb, ub, raw)
        best = y if best is None or y < best else best
        raws.append(raw); ys.append(y); best_list.append(best)
        inp = torch.cat((h, torch.tensor([y], dtype=torch.float64), torch.tensor([1.0], dtype=torch.float64)))
        if (optimum is not None and y <= 1e-8) or len(ys) >= T_MAX:
            break
    return dict(raw=np.array(raws), y=np.array(ys), cost=best_list[::2], fes=len(ys), ret=float(np.sum(ys) / ys[0]))
