"""GPU: landscape-feature kernel vs the golden outputs of the reference's own functions and the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def test_landscape_features_match_reference_and_oracle():
    from metabox_b200 import landscape as ml
    from oracle import landscape as ol
    G = np.load(os.path.join(GOLDEN, 'landscape.npz'))
    by_shape = {}
    for key in G['cases']:
        by_shape.setdefault(G[f'{key}/x'].shape, []).append(str(key))
    n = 0
    for shape, keys in by_shape.items():
        X = np.stack([G[f'{k}/x'] for k in keys]); C_ = np.stack([G[f'{k}/c'] for k in keys]); R = np.stack([G[f'{k}/ref'] for k in keys])
        F = ml.features(X, C_, R).cpu().numpy()
        F0 = ml.features(X, C_, None).cpu().numpy()
        for j, k in enumerate(keys):
            np.testing.assert_allclose(F[j, :4], G[f'{k}/dedqn'], rtol=1e-9, atol=1e-12, err_msg=k)      # the reference itself
            want = ol.features(X[j], C_[j], R[j])
            np.testing.assert_allclose(F[j], want, rtol=1e-9, atol=1e-12, err_msg=k)
            np.testing.assert_allclose(F0[j], ol.features(X[j], C_[j], None), rtol=1e-9, atol=1e-12, err_msg=k)
            a, b = G[f'{k}/hpsde_bool']
            if a >= 0 and abs(F[j, 4] - 0.15) > 1e-9:
                assert int(0.15 < F[j, 4] <= 1) == a, k
            if b >= 0 and abs(F[j, 5] - 0.5) > 1e-9:
                assert int(0.5 <= F[j, 5] <= 1) == b, k
            n += 1
    assert n == len(G['cases'])
    # single-walk helpers with the reference's names; shape errors
    k = str(G['cases'][0])
    x, c = G[f'{k}/x'], G[f'{k}/c']
    assert abs(ml.cal_fdc(x, c) - G[f'{k}/dedqn'][0]) <= 1e-9 * abs(G[f'{k}/dedqn'][0]) + 1e-12
    assert abs(ml.cal_rie(c) - G[f'{k}/dedqn'][1]) <= 1e-9 and abs(ml.cal_acf(c) - G[f'{k}/dedqn'][2]) <= 1e-9
    assert abs(ml.cal_nop(x, c) - G[f'{k}/dedqn'][3]) <= 1e-12
    with pytest.raises(ValueError):
        ml.features(np.zeros((2, 5, 3)), np.zeros((2, 4)))
    with pytest.raises(Exception):
        ml.features(np.zeros((1, 2, 3)), np.zeros((1, 2)))              # S < 3


def test_landscape_features_at_scale():
    """4096 walks of 100 x 30 in one launch: finite, bounded statistics; rie / drie in [0, 1], |acf|, |dfdc| <= 1."""
    import torch
    from metabox_b200 import landscape as ml
    g = torch.Generator(device='cpu').manual_seed(0)
    X = (torch.rand(4096, 100, 30, generator=g, dtype=torch.float64) * 10 - 5).cuda()
    C_ = (X ** 2).sum(-1) + 10 * torch.cos(3 * X).sum(-1)
    F = ml.features(X, C_, None)
    assert torch.all(torch.isfinite(F))
    assert torch.all((F[:, 1] >= 0) & (F[:, 1] <= 1 + 1e-12)) and torch.all((F[:, 5] >= 0) & (F[:, 5] <= 1 + 1e-12))
    assert torch.all(F[:, 2].abs() <= 1 + 1e-9) and torch.all(F[:, 4].abs() <= 1 + 1e-9)
    assert torch.all((F[:, 3] >= 0) & (F[:, 3] < 1))
