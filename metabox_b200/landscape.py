"""Landscape features of sampled walks on the GPU: fitness-distance correlation, ruggedness entropy,
autocorrelation and the number-of-optima proxy of the DEDQN backbone (`src/optimizer/dedqn_optimizer.py:8-76`) and the
dynamic FDC / ruggedness statistics of RL_HPSDE (`src/optimizer/rl_hpsde_optimizer.py:169-216`), batched over walks.

    features(sample [B,S,D], cost [B,S], ref [B,D] | None) -> [B,6] = fdc, rie, acf, nop, dfdc_r, drie_r

The single-walk helpers keep the reference's names and signatures (numpy in, float out)."""
from __future__ import annotations

import numpy as np

from . import _lib

FDC, RIE, ACF, NOP, DFDC_R, DRIE_R = range(6)


def features(sample, cost, ref=None, device=None):
    """torch / numpy inputs; returns a torch float64 tensor [B,6] on the device (no sync)."""
    import torch
    dev = torch.device(device) if device is not None else (sample.device if isinstance(sample, torch.Tensor) and sample.is_cuda
                                                            else torch.device('cuda', torch.cuda.current_device()))
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(device=dev, dtype=torch.float64).contiguous()
    x, c, r = t(sample), t(cost), t(ref)
    if x.dim() != 3 or c.shape != x.shape[:2] or (r is not None and r.shape != (x.shape[0], x.shape[2])):
        raise ValueError(f'sample [B,S,D], cost [B,S], ref [B,D]: got {tuple(x.shape)}, {tuple(c.shape)}, {None if r is None else tuple(r.shape)}')
    B, S, D = x.shape
    out = torch.empty(B, 6, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().mbx_landscape_features(_lib.ptr(x), _lib.ptr(c), _lib.ptr(r), B, S, D, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, 'mbx_landscape_features')
    return out


def _one(sample, fitness, ref=None):
    return features(np.asarray(sample)[None], np.asarray(fitness)[None], None if ref is None else np.asarray(ref)[None])[0].cpu().numpy()


def cal_fdc(sample, fitness):
    return float(_one(sample, fitness)[FDC])


def cal_rie(fitness):
    f = np.asarray(fitness, dtype=np.float64)
    return float(_one(np.zeros((len(f), 1)), f)[RIE])


def cal_acf(fitness):
    f = np.asarray(fitness, dtype=np.float64)
    return float(_one(np.zeros((len(f), 1)), f)[ACF])


def cal_nop(sample, fitness):
    return float(_one(sample, fitness)[NOP])
