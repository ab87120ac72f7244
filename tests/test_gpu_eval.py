"""GPU: problem.eval parity -- CUDA (through the C ABI) vs the golden fixtures and the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _all(dim):
    from metabox_b200 import problem as mp
    out = {}
    for suit in ('bbob', 'bbob-noisy'):
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
        for p in tr.data + te.data:
            out[p.fid] = p
    return out


@pytest.mark.parametrize('dim', [10, 20, 30])
def test_bbob_eval_f64_matches_reference(dim):
    """float64 in / float64 out: |f - f_ref| <= 1e-9 rel (observed ~1e-14), cost within 1e-6 rel + 64 ulp(bias)."""
    from metabox_b200 import problem as mp
    G = np.load(os.path.join(GOLDEN, f'bbob_d{dim}.npz'))
    probs = _all(dim)
    pack = mp.InstancePack([probs[f] for f in sorted(probs)])
    fids = sorted(probs)
    X = np.stack([G[f'x_{f}'].astype(np.float64) for f in fids])
    F = pack.eval_lanes(np.arange(len(fids)), X, noise_free=True)
    worst = 0.
    for k, f in enumerate(fids):
        ref = G[f'f_{f}']
        np.testing.assert_allclose(F[k], ref, rtol=1e-9, err_msg=f'fid {f}')
        opt_ref = float(G[f'opt_{f}'])
        assert abs(probs[f].optimum - opt_ref) <= 4 * np.spacing(abs(opt_ref)), f
        atol = 64 * np.spacing(abs(probs[f].bias) + 1.0)
        np.testing.assert_allclose(F[k] - probs[f].optimum, ref - opt_ref, rtol=1e-6, atol=atol, err_msg=f'fid {f}')
        worst = max(worst, np.max(np.abs(F[k] - ref) / np.abs(ref)))
    print('worst rel', worst)
    assert worst < 1e-11


@pytest.mark.parametrize('dim', [10, 20, 30])
def test_bbob_eval_f32_io(dim):
    """float32 in / float32 out (the storage contract): equals float32(reference on the same points)."""
    from metabox_b200 import problem as mp
    G = np.load(os.path.join(GOLDEN, f'bbob_d{dim}.npz'))
    probs = _all(dim)
    fids = sorted(probs)
    pack = mp.InstancePack([probs[f] for f in fids])
    X = np.stack([G[f'x_{f}'] for f in fids])
    C = pack.eval_lanes(np.arange(len(fids)), X, subtract_optimum=True, noise_free=True)
    assert C.dtype == np.float32
    for k, f in enumerate(fids):
        ref = (G[f'f_{f}'] - float(G[f'opt_{f}']))
        np.testing.assert_allclose(C[k], ref.astype(np.float32), rtol=1e-6,
                                   atol=64 * np.spacing(abs(probs[f].bias) + 1.0), err_msg=f'fid {f}')


@pytest.mark.parametrize('dim', [10, 20, 30])
def test_noisy_eval_with_replayed_draws(dim):
    """Noise models: feed the reference's own draws (regenerated from the recorded seed) through the tape."""
    from metabox_b200 import problem as mp
    from oracle import bbob as ob
    G = np.load(os.path.join(GOLDEN, f'bbob_d{dim}.npz'))
    probs = _all(dim)
    fids = list(range(101, 131))
    pack = mp.InstancePack([probs[f] for f in fids])
    X = np.stack([G[f'x_{f}'].astype(np.float64) for f in fids])
    n = X.shape[1]
    tape = np.zeros((len(fids), 3, n))
    for k, f in enumerate(fids):
        np.random.seed(777 + f)
        var = (f - 101) % 3
        if var == 0:
            tape[k, 0] = np.random.randn(n)
        elif var == 1:
            tape[k, 0] = np.random.rand(n); tape[k, 1] = np.random.rand(n)
        else:
            tape[k, 0] = np.random.rand(n); tape[k, 1] = np.random.randn(n); tape[k, 2] = np.random.randn(n)
    F = pack.eval_lanes(np.arange(len(fids)), X, noise_tape=tape)
    for k, f in enumerate(fids):
        np.testing.assert_allclose(F[k], G[f'fn_{f}'], rtol=1e-9, err_msg=f'fid {f}')


def test_noisy_eval_philox_matches_oracle():
    """Production noise path: Philox draws in-kernel == the oracle's numpy Philox twin."""
    from metabox_b200 import problem as mp
    from oracle import bbob as ob, rng as orng, rlepso as orl
    dim = 10
    G = np.load(os.path.join(GOLDEN, f'bbob_d{dim}.npz'))
    probs = _all(dim)
    tr, te = ob.make_suite('bbob-noisy', dim)
    oinst = {I.fid: I for I in tr + te}
    fids = [101, 102, 103, 107, 114, 118, 121, 124, 126, 130]
    pack = mp.InstancePack([probs[f] for f in fids])
    X = np.stack([G[f'x_{f}'].astype(np.float64) for f in fids])
    seed, gen = 0x1234567890ABCDEF, 7
    F = pack.eval_lanes(np.arange(len(fids)), X, seed=seed, gen=gen)
    for k, f in enumerate(fids):
        I = oinst[f]
        draws = orl.philox_noise(orng.PhiloxDraws(seed, k, gen), I.noise, X.shape[1], orng.TAG_NOISE)
        ref = ob.apply_noise(I, ob.func_true(I, X[k]), draws=draws)
        np.testing.assert_allclose(F[k], ref, rtol=1e-9, err_msg=f'fid {f}')


def test_protein_energy_matches_reference():
    from metabox_b200 import problem as mp
    G = np.load(os.path.join(GOLDEN, 'protein.npz'))
    probs = [mp.Protein_Docking(G[f'coor_init_{p}'], G[f'q_{p}'], G[f'e_{p}'], G[f'r_{p}'], G[f'basis_{p}'],
                                G[f'eigval_{p}'], str(p)) for p in G['ids']]
    pack = mp.InstancePack(probs)
    X = np.stack([G[f'x_{p}'].astype(np.float64) for p in G['ids']])
    E = pack.eval_lanes(np.arange(len(probs)), X)
    for k, p in enumerate(G['ids']):
        np.testing.assert_allclose(E[k], G[f'energy_{p}'], rtol=1e-9, err_msg=str(p))
    E32 = pack.eval_lanes(np.arange(len(probs)), X.astype(np.float32))
    for k, p in enumerate(G['ids']):
        np.testing.assert_allclose(E32[k], G[f'energy_{p}'].astype(np.float32), rtol=1e-6)


def test_problem_object_api_and_edge_cases():
    from metabox_b200 import problem as mp
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    p = te.data[0]                       # Sphere
    p.reset()
    assert p.T1 == 0
    y1 = p.eval(p.opt)                   # 1-D -> scalar, equals optimum exactly
    assert np.ndim(y1) == 0 and y1 == p.optimum
    Y = p.eval(np.zeros((7, 10)))        # 2-D
    assert Y.shape == (7,)
    Y3 = p.eval(np.zeros((2, 3, 10)))    # N-D flattened like the reference
    assert Y3.shape == (6,)
    assert p.T1 > 0
    pack = p._pack
    empty = pack.eval_lanes(np.zeros(0, dtype=np.int32), np.zeros((0, 5, 10)))
    assert empty.shape == (0, 5)
    big = pack.eval_lanes([0], np.random.RandomState(0).uniform(-5, 5, size=(1, 1000, 10)))   # NP > one CTA, ragged tail
    assert big.shape == (1, 1000) and np.all(np.isfinite(big))
    one = np.stack([p.eval(r) for r in np.random.RandomState(0).uniform(-5, 5, size=(1000, 10))[:5]])
    np.testing.assert_array_equal(one, big[0, :5])


def test_protein_in_bounds_pair_pruning_and_out_of_bounds_fallback():
    """In-bounds candidates use the host's candidate-pair list (pairs that can never come within the 9 A cut-off are
    skipped, exactly); candidates outside [-1.5, 1.5]^12 (reachable through mbx_eval only) take all 4950 pairs.  Both vs
    the oracle's dense [100,100] evaluation."""
    from metabox_b200 import problem as mp
    from oracle import protein as op
    G = np.load(os.path.join(GOLDEN, 'protein.npz'))
    ids = [str(p) for p in G['ids']]
    args = lambda p: [G[f'{k}_{p}'] for k in ('coor_init', 'q', 'e', 'r', 'basis', 'eigval')]
    probs = [mp.Protein_Docking(*args(p), p) for p in ids]
    pack = mp.InstancePack(probs)
    n_pairs = [len(P.candidate_pairs()) for P in probs]
    assert all(0 < n < 4950 for n in n_pairs), n_pairs                      # the list really prunes
    g = np.random.RandomState(9)
    Xin = g.uniform(-1.5, 1.5, (len(ids), 24, 12)); Xin[:, 0] = 1.5; Xin[:, 1] = -1.5     # incl. the corners of the box
    Xout = g.uniform(-4.0, 4.0, (len(ids), 24, 12)); Xout[:, 0, 3] = 1.5000001
    for X in (Xin, Xout):
        E = pack.eval_lanes(np.arange(len(ids)), X)
        for k, p in enumerate(ids):
            I = op.ProteinInstance(p, *args(p))
            np.testing.assert_allclose(E[k], op.energy(I, X[k]), rtol=1e-9, atol=1e-9, err_msg=p)
