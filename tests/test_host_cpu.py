"""CPU: host-side logic of the product (instance construction, packing, C-ABI surface).  No compute calls."""
import argparse
import ctypes
import os
import re

import numpy as np
import pytest

from metabox_b200 import _lib, problem as mp

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'metabox_b200.h')).read()
    declared = set(re.findall(r'\b(mbx_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f'{name} declared in include/metabox_b200.h but not exported'
    bound = set(_lib.SIGNATURES) | set(_lib.OPTIONAL_SIGNATURES)
    assert declared <= bound, f'symbols without a ctypes signature: {declared - bound}'
    lib = _lib.lib()
    assert lib.mbx_version() >= 100
    assert lib.mbx_sizeof_instance() == _lib.INSTANCE_DTYPE.itemsize


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    lib = _lib.lib()
    rc = lib.mbx_eval(0, None, None, None, None, 0, None, 1, 1, 10, 0, 0, 0, None, None)
    assert rc == -1 and b'null' in lib.mbx_last_error()
    cfg = _lib.MbxRunConfig(kind=0, NP=3, dim=10, maxFEs=100, n_logpoint=50, log_interval=2, use_tape=0, seed=1)
    rc = lib.mbx_rlepso_init(None, None, None, ctypes.byref(cfg), None, None, None, 1, None)
    assert rc == -1 and b'NP' in lib.mbx_last_error()
    n = lib.mbx_gleet_policy_weight_count()
    assert n == 5432
    assert lib.mbx_gleet_policy(None, None, n, None, None, None, None, None, 1, 129, None) == -1 and b'NP' in lib.mbx_last_error()
    assert lib.mbx_gleet_policy(None, None, n - 1, None, None, None, None, None, 1, 100, None) == -1 and b'packed weights' in lib.mbx_last_error()
    assert lib.mbx_gleet_policy(None, None, n, None, None, None, None, None, 1, 100, None) == -1 and b'null' in lib.mbx_last_error()
    assert lib.mbx_gleet_policy(ctypes.c_void_p(256), ctypes.c_void_p(260), n, None, ctypes.c_void_p(256), None, None, None, 1, 100, None) == -1 \
        and b'aligned' in lib.mbx_last_error()
    assert lib.mbx_gleet_policy(None, None, n, None, None, None, None, None, 0, 100, None) == 0          # empty batch: nothing to do
    k = lib.mbx_rlepso_policy_weight_count()
    assert k == 6726
    assert lib.mbx_rlepso_policy(None, None, k - 2, None, None, None, None, 1, None) == -1 and b'packed weights' in lib.mbx_last_error()
    assert lib.mbx_rlepso_policy(None, None, k, None, None, None, None, 1, None) == -1 and b'null' in lib.mbx_last_error()
    assert lib.mbx_rlepso_policy(None, None, k, None, None, None, None, 0, None) == 0
    assert lib.mbx_rl_pso_policy_weight_count(20) == 2 * (20 * 32 + 32 + 8 * 32 + 8 + 8 + 1)
    assert lib.mbx_rl_pso_policy(None, None, 10, None, None, None, None, 1, 20, None) == -1 and b'packed weights' in lib.mbx_last_error()
    assert lib.mbx_rl_pso_policy(None, None, 10, None, None, None, None, 1, 65, None) == -1 and b'n_in' in lib.mbx_last_error()
    assert lib.mbx_qlpso_policy(None, None, None, None, 1, None) == -1 and b'null' in lib.mbx_last_error()
    assert lib.mbx_qlpso_policy(None, None, None, None, 0, None) == 0
    assert lib.mbx_progressive_walk(None, None, None, None, -5., 5., 10., 1, 10, 200, None, None, None) == -1 and b'null' in lib.mbx_last_error()
    assert lib.mbx_progressive_walk(None, None, None, None, 5., -5., 10., 1, 10, 200, None, None, None) == -1 and b'bad arguments' in lib.mbx_last_error()
    assert lib.mbx_progressive_walk(None, None, None, None, -5., 5., 10., 0, 10, 200, None, None, None) == 0
    m = lib.mbx_deddqn_policy_weight_count()
    assert m == 40704
    assert lib.mbx_deddqn_policy(None, None, m + 1, None, None, 1, None) == -1 and b'packed weights' in lib.mbx_last_error()
    assert lib.mbx_deddqn_policy(None, None, m, None, None, 1, None) == -1 and b'null' in lib.mbx_last_error()
    assert lib.mbx_deddqn_policy(None, None, m, None, None, 0, None) == 0


@pytest.mark.parametrize('dim', [10])
def test_dataset_matches_reference_constants(dim):
    G = np.load(os.path.join(GOLDEN, f'bbob_d{dim}.npz'))
    for suit in ('bbob', 'bbob-noisy'):
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5, difficulty='easy')
        assert [str(p) for p in te.data] == list(G[f'split_easy_{suit}'])
        for p in tr.data + te.data:
            fid = p.fid
            assert str(G[f'name_{fid}']) == str(p)
            np.testing.assert_array_equal(G[f'shift_{fid}'], p.shift)
            np.testing.assert_array_equal(G[f'R_{fid}'], p.rotate)
            assert float(G[f'bias_{fid}']) == p.bias
            if f'M2_{fid}' in G.files:
                np.testing.assert_array_equal(G[f'M2_{fid}'], p.M2)
            if f'y_{fid}' in G.files:
                for k, a in zip(('y', 'C', 'w'), p.peaks):
                    np.testing.assert_array_equal(G[f'{k}_{fid}'], a)
        _, ted = mp.BBOB_Dataset.get_datasets(suit, dim, 5, difficulty='difficult')
        assert [str(p) for p in ted.data] == list(G[f'split_difficult_{suit}'])


def test_pack_layout():
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    probs = tr.data + te.data
    H, Cst = mp.pack_host(probs)
    D = 10
    for h, p in zip(H, probs):
        off = int(h['off'])
        np.testing.assert_array_equal(Cst[off:off + D], p.shift)
        np.testing.assert_array_equal(Cst[off + D:off + D + D * D].reshape(D, D), p.rotate)
        assert h['base'] == p.base and h['dim'] == D and h['fid'] == p.fid
        if p.peaks is not None:
            o = off + 4 * D + 2 * D * D
            np.testing.assert_array_equal(Cst[o:o + p.peaks[0].size].reshape(-1, D), p.peaks[0])
    assert Cst.dtype == np.float64


def test_protein_split_matches_reference():
    G = np.load(os.path.join(GOLDEN, 'protein.npz'))
    tr, te = mp.Protein_Docking_Dataset.split('easy')
    assert [f'{n}_{j + 1}' for n in tr for j in range(10)] == list(G['split_easy_train'])
    assert [f'{n}_{j + 1}' for n in te for j in range(10)] == list(G['split_easy_test'])


def test_dataset_container_api():
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5, train_batch_size=4)
    assert len(tr) == 18 and len(te) == 6 and len(tr + te) == 24
    assert len(tr[0]) == 4
    with pytest.raises(ValueError):
        mp.BBOB_Dataset.get_datasets('cec', 10, 5)
    with pytest.raises(ValueError):
        mp.BBOB_Dataset.get_datasets('bbob', 10, 5, difficulty='hard')


def test_no_cpu_fallback_for_eval():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    tr, _ = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    with pytest.raises(_lib.MbxError):
        tr.data[0].eval(np.zeros(10))


# ----------------------------------------------------------------------------------------- agents (rollout forward)
def test_policy_forward_matches_reference_kats():
    """Weights + known-answer forwards extracted from the reference's shipped pickles (tests/golden/agents.npz)."""
    import torch
    from metabox_b200 import agent as ma
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    sub = lambda pre: {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), sub('rlepso/'))
    mu, sg = actor.distribution(torch.as_tensor(G['rlepso_kat/x']))
    np.testing.assert_allclose(mu.detach().numpy(), G['rlepso_kat/mu'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sg.detach().numpy(), G['rlepso_kat/sigma'], rtol=1e-5, atol=1e-6)
    a = actor(torch.as_tensor(G['rlepso_kat/x']), noise=torch.zeros(5, 35))
    assert a.shape == (5, 35) and float(a.min()) >= 0 and float(a.max()) <= 1
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), sub('deddqn/'))
    q = net(torch.as_tensor(G['deddqn_kat/x']))
    np.testing.assert_allclose(q.detach().numpy(), G['deddqn_kat/q'], rtol=1e-4, atol=1e-5)
    assert sum(p.numel() for p in net.parameters()) == 99 * 100 + 100 + 3 * (100 * 100 + 100) + 100 * 4 + 4
    lde = ma.load_reference_weights(ma.LDE_PolicyNet(), sub('lde/'))
    x, h, c = (torch.as_tensor(G[f'lde_kat/{k}']) for k in ('x', 'h', 'c'))
    mu, sigma, h2, c2 = lde.forward(x, h, c)                                 # nn.LSTM path (reference layout [1,B,*])
    for got, key in ((mu, 'mu'), (sigma, 'sigma'), (h2, 'h_out'), (c2, 'c_out')):
        np.testing.assert_allclose(got.detach().numpy(), G[f'lde_kat/{key}'], rtol=1e-5, atol=1e-6)
    mu, sigma, h2, c2 = lde.lstm_cell(x[0], h[0], c[0])                           # written-out cell used by the sweep
    for got, key in ((mu, 'mu'), (sigma, 'sigma'), (h2, 'h_out'), (c2, 'c_out')):
        np.testing.assert_allclose(got.detach().numpy(), G[f'lde_kat/{key}'][0], rtol=1e-5, atol=1e-6)
    act, _, _ = lde.step(x[0], h[0], c[0], noise=torch.zeros(3, 100))
    np.testing.assert_allclose(act.numpy(), np.clip(G['lde_kat/mu'][0], 0, 1), rtol=1e-5, atol=1e-6)


def test_gleet_and_rl_pso_policies_match_reference_kats():
    """GLEET attention actor / critic (checkpoints 20 and 0 of the shipped rollout agents) and the RL_PSO policy network: weights +
    forwards of the unmodified reference (tests/golden/agents2.npz, made by make_golden_agents2.py).  Case 'r' = states the reference
    GLEET_Optimizer produced; checkpoint 0 keeps mu / sigma away from saturation, so the heads are pinned too."""
    import torch
    from metabox_b200 import agent as ma
    G = np.load(os.path.join(GOLDEN, 'agents2.npz'))
    sub = lambda pre: {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    for grp in ('gleet', 'gleet0'):
        actor = ma.load_reference_weights(ma.GLEET_Actor(), sub(grp + '/'))
        critic = ma.load_reference_weights(ma.GLEET_Critic(), sub(grp + '_critic/'))
        assert sum(p.numel() for p in actor.parameters()) == sum(v.size for v in sub(grp + '/').values())
        assert (actor.max_sigma, actor.min_sigma) == (float(G[grp + '_kat/max_sigma']), float(G[grp + '_kat/min_sigma']))
        for name in ('a', 'b', 'r'):                                   # [3, 100, 27], a ragged swarm [2, 7, 27], reference states [4, 100, 27]
            K = {k: torch.as_tensor(G[f'{grp}_kat/{name}/{k}']) for k in ('x', 'logits', 'mu', 'sigma', 'fixed', 'logp', 'value')}
            mu, sigma, lg = actor.distribution(K['x'])
            np.testing.assert_allclose(lg.detach().numpy(), K['logits'].numpy(), rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(mu.detach().numpy(), K['mu'].numpy()[..., 0], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(sigma.detach().numpy(), K['sigma'].numpy()[..., 0], rtol=1e-4, atol=1e-5)
            lp = actor.log_prob(K['x'], K['fixed'][..., 0])
            np.testing.assert_allclose(lp.detach().numpy(), K['logp'].numpy().reshape(-1), rtol=1e-4, atol=1e-3)
            np.testing.assert_allclose(critic(lg).detach().numpy(), K['value'].numpy().reshape(-1), rtol=1e-4, atol=1e-5)
            a = actor(K['x'], noise=torch.zeros(K['x'].shape[:2]))
            np.testing.assert_allclose(a.numpy(), np.clip(K['mu'].numpy()[..., 0], 0, 1), rtol=1e-4, atol=1e-5)
            assert float(actor(K['x']).min()) >= 0 and float(actor(K['x']).max()) <= 1
    assert float(np.ptp(G['gleet0_kat/r/mu'])) > 0.1 and float(G['gleet0_kat/r/sigma'].max()) < 0.6       # unsaturated heads
    # lanes are independent: a lane's action does not depend on its batch neighbours
    x = torch.as_tensor(G['gleet_kat/a/x'])
    np.testing.assert_allclose(actor.distribution(x[1:2])[0].detach().numpy(), actor.distribution(x)[0][1:2].detach().numpy(), rtol=1e-5, atol=1e-6)
    W = sub('rl_pso/')
    net = ma.load_reference_weights(ma.RL_PSO_PolicyNet(W['_PolicyNetwork__mu_net.net.layer0-linear.weight'].shape[1] // 2), W)
    assert (net.max_sigma, net.min_sigma) == (float(G['rl_pso_kat/max_sigma']), float(G['rl_pso_kat/min_sigma']))
    x = torch.as_tensor(G['rl_pso_kat/x'])
    mu, sigma = net.distribution(x)
    np.testing.assert_allclose(mu.detach().numpy(), G['rl_pso_kat/mu'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sigma.detach().numpy(), G['rl_pso_kat/sigma'], rtol=1e-5, atol=1e-6)
    a = net(x, noise=torch.as_tensor(G['rl_pso_kat/eps']))
    np.testing.assert_allclose(a.numpy(), G['rl_pso_kat/action'], rtol=1e-5, atol=1e-6)          # incl. 3 out-of-range samples
    assert a.shape == (9, 1)
    dq = ma.load_reference_weights(ma.DEDQN_Net(), sub('dedqn/'))                      # Q-network of the shipped DEDQN agent, states of reference runs
    q = dq(torch.as_tensor(G['dedqn_kat/x']))
    np.testing.assert_allclose(q.detach().numpy(), G['dedqn_kat/q'], rtol=1e-5, atol=1e-6)
    assert torch.equal(dq.act(torch.as_tensor(G['dedqn_kat/x'])), torch.as_tensor(np.argmax(G['dedqn_kat/q'], axis=1)))
    for key in ('qlpso_q', 'rl_hpsde_q'):                                              # the shipped tabular agents load into the table policy
        qt = ma.QLPSO_QTable(G[key])
        a = qt(torch.arange(4, dtype=torch.int32), u=torch.full((4,), 0.5))
        want = [int(np.searchsorted(np.cumsum(np.exp(G[key][s]) / np.exp(G[key][s]).sum()), 0.5, side='right')) for s in range(4)]
        assert a.tolist() == [min(w, 3) for w in want] and G[key].shape == (4, 4)


def _emul_lib(name='gleet_policy', defines=()):
    """g++ build of tests/emul/<name>_host.cpp (the device header compiled for the host); built once per session in a temp dir."""
    import ctypes as C
    import subprocess
    import tempfile
    cache = _emul_lib.__dict__.setdefault('cache', {})
    key = (name,) + tuple(defines)
    if key not in cache:
        d = tempfile.mkdtemp(prefix='mbx_emul_')
        so = os.path.join(d, name + '_host.so')
        src = os.path.join(os.path.dirname(__file__), 'emul', name + '_host.cpp')
        subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off', *defines, src, '-o', so], check=True)
        L = C.CDLL(so)
        fp = C.POINTER(C.c_float)
        if name == 'gleet_policy':
            L.glp_host_forward.restype = C.c_int
            L.glp_host_forward.argtypes = [fp, fp, fp, fp, fp, fp, fp, C.c_int, C.c_int]
        elif name == 'pso_policies':
            L.rpp_host_forward.restype = C.c_int
            L.rpp_host_forward.argtypes = [fp, fp, fp, fp, fp, fp, C.c_int, C.c_int]
            L.qlp_host_forward.restype = C.c_int
            L.qlp_host_forward.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_double), fp, C.POINTER(C.c_int32), C.c_int]
        elif name == 'rlepso_policy':
            L.rlp_host_forward.restype = C.c_int
            L.rlp_host_forward.argtypes = [fp, fp, fp, fp, fp, fp, C.c_int]
        else:
            L.ddq_host_forward.restype = C.c_int
            L.ddq_host_forward.argtypes = [fp, fp, C.POINTER(C.c_int32), fp, C.c_int]
        cache[key] = L
    return cache[key]


def _emul_forward(w, x, noise=None, defines=()):
    import ctypes as C
    L = _emul_lib('gleet_policy', defines)
    fp = C.POINTER(C.c_float)
    x = np.ascontiguousarray(x, dtype=np.float32); w = np.ascontiguousarray(w, dtype=np.float32)
    B, NP, _ = x.shape
    act = np.empty((B, NP), np.float32); mu = np.empty_like(act); sg = np.empty_like(act); lg = np.empty((B, NP, 16), np.float32)
    nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
    rc = L.glp_host_forward(x.ctypes.data_as(fp), w.ctypes.data_as(fp), None if nz is None else nz.ctypes.data_as(fp), act.ctypes.data_as(fp),
                            mu.ctypes.data_as(fp), sg.ctypes.data_as(fp), lg.ctypes.data_as(fp), B, NP)
    assert rc == 0
    return act, mu, sg, lg


def test_fused_gleet_policy_arithmetic_on_the_host_matches_the_reference():
    """The per-particle device functions of the fused GLEET policy kernel (csrc/mbx_gleet_policy.cuh) compiled for the host, driven in the
    kernel's phase order, with the weights packed by GLEET_Actor.pack(): against forwards of the unmodified reference (both checkpoints,
    random / ragged / reference-produced states) and against the torch module under action noise."""
    import torch
    from metabox_b200 import agent as ma
    G = np.load(os.path.join(GOLDEN, 'agents2.npz'))
    sub = lambda pre: {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    assert _emul_lib().glp_host_weight_count() == 5432
    g = np.random.RandomState(2)
    for grp in ('gleet', 'gleet0'):
        actor = ma.load_reference_weights(ma.GLEET_Actor(), sub(grp + '/'))
        w = actor.pack().numpy()
        assert w.shape == (5432,) and w.dtype == np.float32
        for name in ('a', 'b', 'r'):
            x = G[f'{grp}_kat/{name}/x']
            noise = g.randn(*x.shape[:2]).astype(np.float32)
            act, mu, sg, lg = _emul_forward(w, x, noise)
            np.testing.assert_allclose(lg, G[f'{grp}_kat/{name}/logits'], rtol=2e-4, atol=3e-5)
            np.testing.assert_allclose(mu, G[f'{grp}_kat/{name}/mu'][..., 0], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(sg, G[f'{grp}_kat/{name}/sigma'][..., 0], rtol=1e-4, atol=1e-5)
            want = actor(torch.as_tensor(x), noise=torch.as_tensor(noise)).numpy()
            np.testing.assert_allclose(act, want, rtol=1e-4, atol=2e-5)
            assert act.min() >= 0 and act.max() <= 1
            np.testing.assert_allclose(_emul_forward(w, x)[0], np.clip(mu, 0, 1), rtol=0, atol=0)        # null noise -> the mean action
    # the A/B variant with packed f32x2 arithmetic sums the head dot products in another order: same answers to rounding
    actor = ma.load_reference_weights(ma.GLEET_Actor(), sub('gleet0/'))
    x = G['gleet0_kat/r/x']
    act, mu, sg, lg = _emul_forward(actor.pack().numpy(), x, defines=('-DGLP_PACKED_F32X2',))
    np.testing.assert_allclose(lg, G['gleet0_kat/r/logits'], rtol=2e-4, atol=3e-5)
    np.testing.assert_allclose(mu, G['gleet0_kat/r/mu'][..., 0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(sg, G['gleet0_kat/r/sigma'][..., 0], rtol=1e-4, atol=1e-5)
    # edge swarm sizes (1 particle: the unbiased variance has 15 degrees of freedom; 128 = the kernel's maximum)
    actor = ma.load_reference_weights(ma.GLEET_Actor(), sub('gleet0/'))
    for NP in (1, 2, 37, 128):
        x = g.rand(2, NP, 27).astype(np.float32)
        act, mu, sg, lg = _emul_forward(actor.pack().numpy(), x)
        m2, s2, l2 = actor.distribution(torch.as_tensor(x))
        np.testing.assert_allclose(lg, l2.detach().numpy(), rtol=2e-4, atol=3e-5)
        np.testing.assert_allclose(mu, m2.detach().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sg, s2.detach().numpy(), rtol=1e-4, atol=1e-5)


def test_fused_rlepso_policy_arithmetic_on_the_host_matches_the_reference():
    """The per-neuron device functions of the fused RLEPSO actor kernel (csrc/mbx_rlepso_policy.cuh) compiled for the host with the weights
    packed by RLEPSO_Actor.pack(): (mu, sigma) of the unmodified reference on the shipped pickle (agents.npz) and the torch module's
    sampled actions under the same noise."""
    import ctypes as C
    import torch
    from metabox_b200 import agent as ma
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), {k[len('rlepso/'):]: G[k] for k in G.files if k.startswith('rlepso/')})
    w = np.ascontiguousarray(actor.pack().numpy())
    L = _emul_lib('rlepso_policy')
    assert L.rlp_host_weight_count() == w.size == 2 * (64 + 64 + 64 * 32 + 32 + 32 * 35 + 35)
    fp = C.POINTER(C.c_float)

    def run(x, noise=None):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        act = np.empty((len(x), 35), np.float32); mu = np.empty_like(act); sg = np.empty_like(act)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
        assert L.rlp_host_forward(x.ctypes.data_as(fp), w.ctypes.data_as(fp), None if nz is None else nz.ctypes.data_as(fp),
                                  act.ctypes.data_as(fp), mu.ctypes.data_as(fp), sg.ctypes.data_as(fp), len(x)) == 0
        return act, mu, sg

    act, mu, sg = run(G['rlepso_kat/x'])
    np.testing.assert_allclose(mu, G['rlepso_kat/mu'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sg, G['rlepso_kat/sigma'], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(act, np.clip(mu, 0, 1))
    g = np.random.RandomState(4)
    x = g.rand(37, 1).astype(np.float32); noise = g.randn(37, 35).astype(np.float32)
    act, mu, sg = run(x, noise)
    want = actor(torch.as_tensor(x), noise=torch.as_tensor(noise)).numpy()
    np.testing.assert_allclose(act, want, rtol=1e-5, atol=2e-6)
    act, _, _ = run(x, 50 * noise)                                                  # wild draws: the clamp to [0, 1] is part of the action
    np.testing.assert_allclose(act, actor(torch.as_tensor(x), noise=torch.as_tensor(50 * noise)).numpy(), rtol=1e-5, atol=2e-6)
    assert act.min() == 0 and act.max() == 1


def test_fused_rl_pso_and_qlpso_policy_arithmetic_on_the_host_matches_the_reference():
    """Device functions of the RL_PSO policy-network kernel and the QLPSO table-sampling kernel (csrc/mbx_pso_policies.cuh) compiled for the
    host: RL_PSO against the unmodified reference's (mu, sigma) and its out-of-range transformation (agents2.npz), other input widths
    against the torch module; QLPSO against QLPSO_QTable.forward for every state and a sweep of uniforms."""
    import ctypes as C
    import torch
    from metabox_b200 import agent as ma
    L = _emul_lib('pso_policies')
    fp = C.POINTER(C.c_float)
    G = np.load(os.path.join(GOLDEN, 'agents2.npz'))
    net = ma.load_reference_weights(ma.RL_PSO_PolicyNet(10), {k[len('rl_pso/'):]: G[k] for k in G.files if k.startswith('rl_pso/')})

    def run(net, x, noise=None):
        w = np.ascontiguousarray(net.pack().numpy())
        assert L.rpp_host_weight_count(2 * net.dim) == w.size
        x = np.ascontiguousarray(x, dtype=np.float32)
        act = np.empty(len(x), np.float32); mu = np.empty_like(act); sg = np.empty_like(act)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32).reshape(-1)
        assert L.rpp_host_forward(x.ctypes.data_as(fp), w.ctypes.data_as(fp), None if nz is None else nz.ctypes.data_as(fp),
                                  act.ctypes.data_as(fp), mu.ctypes.data_as(fp), sg.ctypes.data_as(fp), len(x), x.shape[1]) == 0
        return act, mu, sg

    act, mu, sg = run(net, G['rl_pso_kat/x'], G['rl_pso_kat/eps'])
    np.testing.assert_allclose(mu, G['rl_pso_kat/mu'][:, 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sg, G['rl_pso_kat/sigma'][:, 0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(act, G['rl_pso_kat/action'][:, 0], rtol=1e-5, atol=2e-6)        # 4 of the 9 samples take the transformation
    g = np.random.RandomState(6)
    for dim in (12, 30):
        torch.manual_seed(dim)
        n2 = ma.RL_PSO_PolicyNet(dim)
        x = (g.rand(11, 2 * dim) * 2 - 0.5).astype(np.float32); nz = (3 * g.randn(11, 1)).astype(np.float32)
        act, _, _ = run(n2, x, nz)
        np.testing.assert_allclose(act, n2(torch.as_tensor(x), noise=torch.as_tensor(nz)).numpy()[:, 0], rtol=1e-5, atol=2e-6)
    qt = np.array([[0., 1., 2., 3.], [0., 0., 0., 0.], [5., 0., 0., 0.], [-1., -1., 4., -1.]])
    q = ma.QLPSO_QTable(qt)
    n = 4000
    state = np.ascontiguousarray(g.randint(0, 4, size=n), dtype=np.int32)
    u = np.ascontiguousarray(g.rand(n), dtype=np.float32); u[:4] = 0.0; u[4:8] = np.float32(1 - 2 ** -24)
    act = np.empty(n, np.int32)
    assert L.qlp_host_forward(state.ctypes.data_as(C.POINTER(C.c_int32)), qt.ctypes.data_as(C.POINTER(C.c_double)), u.ctypes.data_as(fp),
                              act.ctypes.data_as(C.POINTER(C.c_int32)), n) == 0
    np.testing.assert_array_equal(act, q(torch.as_tensor(state), u=torch.as_tensor(u)).numpy())
    assert set(act.tolist()) == {0, 1, 2, 3}


def test_fused_deddqn_policy_arithmetic_on_the_host_matches_the_reference():
    """The per-neuron device functions of the fused DE-DDQN Q-network kernel (csrc/mbx_deddqn_policy.cuh) compiled for the host, driven in
    the kernel's order with the weights packed by DE_DDQN_Net.pack(): Q-values of the unmodified reference on the shipped pickle
    (agents.npz), greedy actions equal to torch.argmax, ragged tile (B not a multiple of 8), NaN / inf features count as 0."""
    import ctypes as C
    import torch
    from metabox_b200 import agent as ma
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), {k[len('deddqn/'):]: G[k] for k in G.files if k.startswith('deddqn/')})
    w = np.ascontiguousarray(net.pack().numpy())
    L = _emul_lib('deddqn_policy')
    assert L.ddq_host_weight_count() == w.size == 99 * 100 + 100 + 3 * (100 * 100 + 100) + 4 * 100 + 4
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)

    def run(x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        act = np.empty(len(x), np.int32); q = np.empty((len(x), 4), np.float32)
        assert L.ddq_host_forward(x.ctypes.data_as(fp), w.ctypes.data_as(fp), act.ctypes.data_as(ip), q.ctypes.data_as(fp), len(x)) == 0
        return act, q

    act, q = run(G['deddqn_kat/x'])                                    # 16 states
    np.testing.assert_allclose(q, G['deddqn_kat/q'], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(act, np.argmax(G['deddqn_kat/q'], axis=1))
    g = np.random.RandomState(7)
    x = (g.rand(21, 99) * 2 - 0.5).astype(np.float32)                 # 21 lanes: two full tiles + 5
    x[3, 17] = np.nan; x[9, 0] = np.inf; x[20, 98] = -np.inf
    act, q = run(x)
    want = net(torch.nan_to_num(torch.as_tensor(x), nan=0.0, posinf=0.0, neginf=0.0)).detach().numpy()
    np.testing.assert_allclose(q, want, rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(act, np.argmax(want, axis=1))
    assert run(x[:1])[0][0] == act[0] and len(run(x[:0])[0]) == 0


# ----------------------------------------------------------------------------------------- sharding / gather (gloo, world 2)
def test_shard_ranges_cover_the_lane_table():
    from metabox_b200 import sweep
    inst, run = sweep.lane_table(7, 5)
    assert list(inst[:6]) == [0, 0, 0, 0, 0, 1] and list(run[:6]) == [0, 1, 2, 3, 4, 0]
    for n, world in ((35, 1), (35, 2), (35, 4), (35, 8), (3, 8), (0, 2)):
        spans = [sweep.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, n_lanes, q):
    import torch
    import torch.distributed as dist
    from metabox_b200 import sweep
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = sweep.shard_range(n_lanes, rank, world)
        lanes = torch.arange(lo, hi, dtype=torch.float32)
        # a fake per-rank sweep result whose content encodes the global lane id
        res = dict(curve=lanes[:, None] + torch.arange(51, dtype=torch.float32)[None, :] / 100.,
                   fes=(1000 + lanes).to(torch.int32), ret=-lanes)
        rows = sweep.gather_results(res, world)                                   # counts gathered first (two collectives)
        rows1 = sweep.gather_results(res, world, n_max=-(-n_lanes // world))      # ONE collective: the count rides in the payload
        assert torch.equal(rows, rows1)
        q.put((rank, rows1.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_lanes', [10, 7])
def test_gather_results_world2_gloo(n_lanes):
    """world_size-2 gloo run of the sweep's single collective: uneven shards, rows come back in global lane order on
    every rank and convert to the reference's result-dict layout."""
    import socket
    import torch.multiprocessing as tmp
    from metabox_b200 import sweep
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_lanes, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lanes = np.arange(n_lanes, dtype=np.float32)
    for r in range(2):
        rows = got[r]
        assert rows.shape == (n_lanes, 53)
        np.testing.assert_allclose(rows[:, 0], lanes); np.testing.assert_allclose(rows[:, 51], 1000 + lanes)
        np.testing.assert_allclose(rows[:, 52], -lanes); np.testing.assert_allclose(rows[:, 50], lanes + 0.5)
    if n_lanes == 10:
        d = sweep.results_to_reference_dict(got[0], ['A', 'B'], 5, 'RLEPSO_Agent')
        assert d['fes']['B']['RLEPSO_Agent'] == [1005, 1006, 1007, 1008, 1009]
        assert len(d['cost']['A']['RLEPSO_Agent']) == 5 and len(d['cost']['A']['RLEPSO_Agent'][0]) == 51


def test_tester_rows_to_reference_dict_layout():
    """Host-side assembly of test.pkl / rollout.pkl entries (tester.py:123-127, 282-300) from gathered result rows."""
    from metabox_b200 import tester as mt
    rows = np.zeros((6, 53))
    rows[:, :51] = np.linspace(5, 1, 51)[None, :] + np.arange(6)[:, None]
    rows[:, 51] = 2000 + np.arange(6)
    rows[:, 52] = -3.0
    out = mt.rows_to_results(rows, ['A', 'B'], 3, 'RLEPSO_Agent', {'cost': {}, 'fes': {}, 'return': {}})
    assert list(out['cost']) == ['A', 'B'] and out['fes']['B']['RLEPSO_Agent'] == [2003, 2004, 2005]
    assert len(out['cost']['A']['RLEPSO_Agent']) == 3 and len(out['cost']['A']['RLEPSO_Agent'][0]) == 51
    assert out['return']['A']['RLEPSO_Agent'] == [-3.0] * 3 and isinstance(out['fes']['A']['RLEPSO_Agent'][0], int)
    out2 = mt.rows_to_results(rows, ['A', 'B'], 3, 'X', {'cost': {}, 'fes': {}})
    assert 'return' not in out2 and mt.cal_t0(4, 50) > 0


# ----------------------------------------------------------------------------------------- LDE training maths (CPU)
def _reference_lde_loss(net, inputs, actions, hs, cs, rewards, traj_num, gamma=0.99):
    """`LDE_Agent.train_episode` (lde_agent.py:116-129) + `__discounted_norm_rewards` (:68-83) restated with the
    reference's own data layout (trajectory-major flat batches, nn.LSTM forward) for ONE problem."""
    import torch
    T, L = rewards.shape                      # L = traj_num trajectories of equal length T
    flat = lambda a: a.permute(1, 0, 2).reshape(L * T, -1)          # trajectory-major like the reference's append order
    mu, sigma, _, _ = net.forward(flat(inputs)[None], flat(hs)[None], flat(cs)[None])
    logp = torch.sum(torch.distributions.Normal(mu[0], sigma[0]).log_prob(flat(actions) + 1e-8), 1)
    r = rewards.t().reshape(-1).numpy()
    disc = np.zeros_like(r)
    for ep in range(traj_num):
        run = 0.
        for t in reversed(range(T)):
            run = run * gamma + r[ep * T + t]
            disc[ep * T + t] = run
    return -torch.mean(logp * torch.as_tensor(disc, dtype=torch.float32))


def test_lde_reinforce_loss_and_discounting_match_the_reference_formulas():
    import torch
    from metabox_b200 import agent as ma, trainer as mt
    torch.manual_seed(0)
    NP, T, TR, P = 10, 7, 20, 3
    net = ma.LDE_PolicyNet(NP=NP)
    L = P * TR
    inputs, actions = torch.rand(T, L, NP + 10), torch.rand(T, L, 2 * NP)
    hs, cs = torch.randn(T, L, 50) * 0.3, torch.randn(T, L, 50) * 0.3
    rewards, weight = torch.rand(T, L), torch.ones(T, L)
    loss, n = mt.reinforce_loss(net, inputs, actions, hs, cs, rewards, weight, TR)
    want = sum(_reference_lde_loss(net, inputs[:, p * TR:(p + 1) * TR], actions[:, p * TR:(p + 1) * TR], hs[:, p * TR:(p + 1) * TR],
                                   cs[:, p * TR:(p + 1) * TR], rewards[:, p * TR:(p + 1) * TR], TR) for p in range(P))
    assert n == P
    np.testing.assert_allclose(float(loss.detach()), float(want.detach()), rtol=2e-5)
    # gradients agree too (the hand-written LSTM cell == nn.LSTM)
    g1 = torch.autograd.grad(loss, list(net.parameters()), retain_graph=True)
    g2 = torch.autograd.grad(want, list(net.parameters()))
    for a, b in zip(g1, g2):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=2e-4, atol=2e-5)
    # finished steps carry zero weight
    w2 = weight.clone(); w2[4:, :5] = 0
    l2, _ = mt.reinforce_loss(net, inputs, actions, hs, cs, rewards * w2, w2, TR)
    assert np.isfinite(float(l2)) and abs(float(l2) - float(loss)) > 0


def _grad_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from metabox_b200 import trainer as mt
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(3, 2)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))        # "sum over local problems" gradients
        total, (alive,) = mt.allreduce_gradients(lin.parameters(), n_local=rank + 1, extra=[10.0 * (rank + 1)])   # rank 0: 1 problem, rank 1: 2
        assert alive == 30.0                                     # extra scalars ride in the same collective (PPO: lanes still running)
        q.put((rank, total, [p.grad.clone().numpy() for p in lin.parameters()]))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    """ONE all-reduce per learn step: gradient of the mean per-problem loss over all ranks' problems, on every rank."""
    import socket
    import torch.multiprocessing as tmp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, grads in got:
        assert total == 3.0
        for g in grads:
            np.testing.assert_allclose(g, (1.0 + 2.0) / 3.0)


def _reference_ppo_epoch(actor, critic, states, actions, old_logp, rewards, next_state, old_value, gamma=0.999, eps_clip=0.1):
    """Literal per-episode restatement of rlepso_agent.py:196-253 (one lane, python loops over the segment)."""
    import torch
    t_time = states.shape[0]
    logprobs, bl_val = [], []
    for tt in range(t_time):
        mu, sigma = actor.distribution(states[tt])
        logprobs.append(torch.sum(torch.distributions.Normal(mu, sigma).log_prob(actions[tt])))
        bl_val.append(critic(states[tt]).squeeze())
    logprobs, bl_val = torch.stack(logprobs).view(-1), torch.stack(bl_val).view(-1)
    R = critic(next_state).detach().squeeze()
    Reward = []
    for r in list(rewards)[::-1]:
        R = R * gamma + r
        Reward.append(R)
    Reward = torch.stack(Reward[::-1], 0).view(-1)
    ratios = torch.exp(logprobs - old_logp.detach())
    advantages = Reward - bl_val.detach()
    reinforce_loss = -torch.min(ratios * advantages, torch.clamp(ratios, 1 - eps_clip, 1 + eps_clip) * advantages).mean()
    if old_value is None:
        baseline_loss = ((bl_val - Reward) ** 2).mean()
    else:
        vpredclipped = old_value + torch.clamp(bl_val - old_value, -eps_clip, eps_clip)
        baseline_loss = torch.max((bl_val - Reward) ** 2, (vpredclipped - Reward) ** 2).mean()
    return reinforce_loss, baseline_loss, bl_val.detach()


def test_rlepso_ppo_losses_match_the_reference_formulas():
    import torch
    from metabox_b200 import agent as ma, trainer as mt
    torch.manual_seed(1)
    actor, critic = ma.RLEPSO_Actor(), ma.RLEPSO_Critic()
    T, L = 10, 5
    states, actions = torch.rand(T, L, 1), torch.rand(T, L, 35)
    rewards, nxt = torch.sign(torch.randn(T, L)), torch.rand(L, 1)
    with torch.no_grad():
        old_logp = actor.log_prob(states, actions) + 0.3 * torch.randn(T, L)         # a policy that has already moved
    weight = torch.ones(T, L)
    weight[6:, 2] = 0                                                                 # lane 2 finished after 6 generations
    for old_value in (None, torch.rand(T, L)):
        rl, bl, val, n = mt.ppo_losses(actor, critic, states, actions, old_logp, rewards * weight, weight, nxt, old_value)
        want_r, want_b = 0., 0.
        for b in range(L):
            tt = 6 if b == 2 else T
            r, v, _ = _reference_ppo_epoch(actor, critic, states[:tt, b], actions[:tt, b], old_logp[:tt, b], rewards[:tt, b], nxt[b],
                                           None if old_value is None else old_value[:tt, b])
            want_r, want_b = want_r + r, want_b + v
        assert n == L
        np.testing.assert_allclose(float(rl.detach()), float(want_r.detach()), rtol=2e-5)
        np.testing.assert_allclose(float(bl.detach()), float(want_b.detach()), rtol=2e-5)
        ps = list(actor.parameters()) + list(critic.parameters())
        g1 = torch.autograd.grad(rl + bl, ps, retain_graph=True, allow_unused=True)
        g2 = torch.autograd.grad(want_r + want_b, ps, allow_unused=True)
        for a, b in zip(g1, g2):
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=3e-4, atol=3e-6)
    # a lane without live steps contributes nothing
    w0 = weight.clone(); w0[:, 4] = 0
    _, _, _, n = mt.ppo_losses(actor, critic, states, actions, old_logp, rewards * w0, w0, nxt)
    assert n == L - 1


class _FakeRlepsoEngine:
    """CPU stand-in with the engine surface RLEPSO_Trainer uses: lane k of the GLOBAL lane table finishes after 4 + 9 k
    generations, rewards are +-1 (0 once finished), the feature is the consumed budget."""

    def __init__(self, pack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, lane_base=0):
        import torch
        self.B = len(lane_inst)
        self.life = torch.tensor([4 + 9 * (lane_base + k) for k in range(self.B)], dtype=torch.int64)
        self.gen = torch.zeros(self.B, dtype=torch.int64)
        self.feature = torch.zeros(self.B)
        self.done = torch.zeros(self.B, dtype=torch.uint8)
        self.g = torch.Generator().manual_seed(seed + lane_base)

    def init_population(self):
        return self.feature

    def update(self, action):
        import torch
        assert action.shape == (self.B, 35)
        alive = self.done == 0
        self.gen += alive.long()
        r = torch.where(torch.rand(self.B, generator=self.g) < action.mean(1), 1.0, -1.0) * alive
        self.done = (self.gen >= self.life).to(torch.uint8)
        self.feature = self.gen.float() / 40.0
        return self.feature, r, self.done

    def cost_curves(self, pad=False):
        return [[10.0, 1.0] for _ in range(self.B)]


def _ppo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from metabox_b200 import problem as mp, trainer as mt
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mp.InstancePack = lambda problems, device=None: None            # no GPU in this test: host logic only
        mt.RLEPSO_Engine = _FakeRlepsoEngine
        torch.manual_seed(0)                                             # identical initial weights on every rank
        cfg = argparse.Namespace(dim=10, maxFEs=4000, n_logpoint=50, log_interval=80)
        T = mt.RLEPSO_Trainer(cfg, [None] * 5, seed=1, device='cpu')
        info = T.train_episodes([0, 1, 2, 3, 4])                         # rank 0: lanes 0-2 (<= 22 gens), rank 1: lanes 3, 4 (<= 40 gens)
        flat = torch.cat([p.detach().reshape(-1) for p in T._params()])
        q.put((rank, info['segments'], info['learn_steps'], len(info['gbest']), flat.numpy()))
    finally:
        dist.destroy_process_group()


def test_rlepso_ppo_loop_world2_gloo_ranks_stay_in_step():
    """Ranks whose lanes finish in different segments still make the same number of collectives and identical updates."""
    import socket
    import torch.multiprocessing as tmp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ppo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=90) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, seg0, ls0, n0, w0), (r1, seg1, ls1, n1, w1) = got
    assert seg0 == seg1 == 4 and ls0 == ls1 == 12                        # longest lane: 4 + 9*4 = 40 generations = 4 segments x 3 epochs
    assert (n0, n1) == (3, 2)
    np.testing.assert_array_equal(w0, w1)                                # the same all-reduced gradients -> the same weights


def _reference_gleet_ppo_epoch(actor, critic, states, actions, old_logp, rewards, next_state, old_value, gamma=0.999, eps_clip=0.1):
    """Literal per-episode restatement of gleet_agent.py:196-262 (one lane, python loops over the segment): the critic reads the actor's
    decoder output, the log-probability is the sum over the swarm, loss = baseline_loss + reinforce_loss."""
    import torch
    logprobs, bl_val = [], []
    for tt in range(states.shape[0]):
        mu, sigma, to_critic = actor.distribution(states[tt][None])
        logprobs.append(torch.sum(torch.distributions.Normal(mu, sigma).log_prob(actions[tt][None]), dim=1))
        bl_val.append(critic(to_critic).squeeze())
    logprobs, bl_val = torch.stack(logprobs).view(-1), torch.stack(bl_val).view(-1)
    R = critic(actor.logits(next_state[None])).detach().squeeze()
    Reward = []
    for r in list(rewards)[::-1]:
        R = R * gamma + r
        Reward.append(R)
    Reward = torch.stack(Reward[::-1], 0).view(-1)
    ratios = torch.exp(logprobs - old_logp.detach())
    advantages = Reward - bl_val.detach()
    reinforce_loss = -torch.min(ratios * advantages, torch.clamp(ratios, 1 - eps_clip, 1 + eps_clip) * advantages).mean()
    if old_value is None:
        baseline_loss = ((bl_val - Reward) ** 2).mean()
    else:
        vpredclipped = old_value + torch.clamp(bl_val - old_value, -eps_clip, eps_clip)
        baseline_loss = torch.max((bl_val - Reward) ** 2, (vpredclipped - Reward) ** 2).mean()
    return reinforce_loss, baseline_loss


def test_gleet_ppo_losses_match_the_reference_formulas():
    """GLEET's PPO epoch through the shared `ppo_losses` (actor / critic adapters of trainer.py) against the literal restatement, lane by
    lane: losses and the gradients of ALL parameters (the baseline loss reaches the actor's attention stack through the critic's input)."""
    import torch
    from metabox_b200 import agent as ma, trainer as mt
    torch.manual_seed(3)
    G = np.load(os.path.join(GOLDEN, 'agents2.npz'))
    sub = lambda pre: {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    actor = ma.load_reference_weights(ma.GLEET_Actor(), sub('gleet0/'))
    critic = ma.load_reference_weights(ma.GLEET_Critic(), sub('gleet0_critic/'))
    T, L, NP = 4, 3, 12
    states, actions = torch.rand(T, L, NP, 27), torch.rand(T, L, NP)
    rewards, nxt = torch.rand(T, L) * 0.2, torch.rand(L, NP, 27)
    ah = mt._GleetActorHead(actor)
    ch = mt._GleetCriticHead(actor, critic, ah)                                       # shares the decoder output of log_prob(states)
    with torch.no_grad():
        old_logp = ah.log_prob(states, actions) + 0.05 * torch.randn(T, L)
    weight = torch.ones(T, L)
    weight[2:, 1] = 0                                                                 # lane 1 finished after 2 generations
    for old_value in (None, torch.rand(T, L)):
        rl, bl, val, n = mt.ppo_losses(ah, ch, states, actions, old_logp, rewards * weight, weight, nxt, old_value)
        want_r, want_b = 0., 0.
        for b in range(L):
            tt = 2 if b == 1 else T
            r, v = _reference_gleet_ppo_epoch(actor, critic, states[:tt, b], actions[:tt, b], old_logp[:tt, b], rewards[:tt, b], nxt[b],
                                              None if old_value is None else old_value[:tt, b])
            want_r, want_b = want_r + r, want_b + v
        assert n == L and val.shape == (T, L)
        np.testing.assert_allclose(float(rl.detach()), float(want_r.detach()), rtol=1e-4)
        np.testing.assert_allclose(float(bl.detach()), float(want_b.detach()), rtol=1e-4)
        ps = list(actor.parameters()) + list(critic.parameters())
        g1 = torch.autograd.grad(rl + bl, ps, retain_graph=True, allow_unused=True)
        g2 = torch.autograd.grad(want_r + want_b, ps, allow_unused=True)
        assert all(a is not None for a in g1)                                         # every parameter of actor AND critic is trained
        for a, b in zip(g1, g2):
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=2e-3, atol=2e-5)


class _FakeGleetEngine:
    """CPU stand-in with the engine surface GLEET_Trainer uses: lane k finishes after 7 + 6 k generations; float rewards."""

    def __init__(self, pack, lane_inst, maxFEs, n_logpoint=50, log_interval=None, seed=0, lane_base=0, NP=6):
        import torch
        self.B, self.NP = len(lane_inst), NP
        self.life = torch.tensor([7 + 6 * (lane_base + k) for k in range(self.B)], dtype=torch.int64)
        self.gen = torch.zeros(self.B, dtype=torch.int64)
        self.g = torch.Generator().manual_seed(seed + lane_base)
        self.feature = torch.rand(self.B, NP, 27, generator=self.g)
        self.done = torch.zeros(self.B, dtype=torch.uint8)

    def init_population(self):
        return self.feature

    def update(self, action):
        import torch
        assert action.shape == (self.B, self.NP) and float(action.min()) >= 0 and float(action.max()) <= 1
        alive = self.done == 0
        self.gen += alive.long()
        r = action.mean(1) * 0.1 * alive
        self.done = (self.gen >= self.life).to(torch.uint8)
        self.feature = torch.rand(self.B, self.NP, 27, generator=self.g)
        return self.feature, r, self.done

    def cost_curves(self, pad=False):
        return [[10.0, 1.0] for _ in range(self.B)]


def test_gleet_ppo_loop_on_cpu_with_a_fake_engine():
    """Host logic of GLEET_Trainer: segments of 10 generations x 3 epochs until the longest lane is done, one Adam optimizer, gradient
    norms clipped to 0.1 per parameter group before every step, parameters move, max_learning_step stops the loop."""
    import torch
    from metabox_b200 import problem as mp, trainer as mt
    old = (mp.InstancePack, mt.GLEET_Engine)
    try:
        mp.InstancePack = lambda problems, device=None: None
        mt.GLEET_Engine = _FakeGleetEngine
        torch.manual_seed(0)
        cfg = argparse.Namespace(dim=10, maxFEs=4000, n_logpoint=50, log_interval=80)
        T = mt.GLEET_Trainer(cfg, [None] * 3, seed=1, device='cpu')
        before = torch.cat([p.detach().reshape(-1).clone() for p in T._params()])
        norms = []
        step = T.optimizer.step
        T.optimizer.step = lambda: (norms.append([float(torch.norm(torch.stack([p.grad.norm() for p in g['params']]))) for g in T.optimizer.param_groups]), step())[1]
        info = T.train_episodes([0, 1, 2])                                    # lanes live 7, 13, 19 generations -> 2 segments
        assert info['segments'] == 2 and info['learn_steps'] == 6 and len(info['gbest']) == 3 and len(norms) == 6
        assert all(n <= 0.1 + 1e-6 for pair in norms for n in pair)
        after = torch.cat([p.detach().reshape(-1) for p in T._params()])
        assert float((after - before).abs().max()) > 0 and torch.isfinite(after).all()
        assert all(r > 0 for r in info['return'])
        cfg.max_learning_step = 8
        info = T.train_episodes([0, 1, 2])
        assert info['done'] and info['learn_steps'] == 8
    finally:
        mp.InstancePack, mt.GLEET_Engine = old


def _gleet_ppo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from metabox_b200 import problem as mp, trainer as mt
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mp.InstancePack = lambda problems, device=None: None            # no GPU in this test: host logic only
        mt.GLEET_Engine = _FakeGleetEngine
        torch.manual_seed(0)                                             # identical initial weights on every rank
        cfg = argparse.Namespace(dim=10, maxFEs=4000, n_logpoint=50, log_interval=80)
        T = mt.GLEET_Trainer(cfg, [None] * 5, seed=1, device='cpu')
        info = T.train_episodes([0, 1, 2, 3, 4])                         # rank 0: lanes 0-2 (<= 19 gens), rank 1: lanes 3, 4 (<= 31 gens)
        flat = torch.cat([p.detach().reshape(-1) for p in T._params()])
        q.put((rank, info['segments'], info['learn_steps'], len(info['gbest']), flat.numpy()))
    finally:
        dist.destroy_process_group()


def test_gleet_ppo_loop_world2_gloo_ranks_stay_in_step():
    """GLEET_Trainer under gloo, world 2: ranks whose lanes finish in different segments make the same number of collectives; the clipped,
    all-reduced gradients give identical weights on both ranks."""
    import socket
    import torch.multiprocessing as tmp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_gleet_ppo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, seg0, ls0, n0, w0), (r1, seg1, ls1, n1, w1) = got
    assert seg0 == seg1 == 4 and ls0 == ls1 == 12                        # longest lane: 7 + 6*4 = 31 generations = 4 segments x 3 epochs
    assert (n0, n1) == (3, 2)
    np.testing.assert_array_equal(w0, w1)


def test_qlpso_qtable_policy_on_cpu():
    """Tabular QLPSO policy (qlpso_agent.py:33-36, 47-48): softmax sampling by inverse CDF, TD rule."""
    import torch
    from metabox_b200 import agent as ma
    qt = np.array([[0., 1., 2., 3.], [0., 0., 0., 0.], [5., 0., 0., 0.], [-1., -1., 4., -1.]])
    q = ma.QLPSO_QTable(qt)
    g = torch.Generator().manual_seed(0)
    n = 100000
    state = torch.randint(0, 4, (n,), generator=g, dtype=torch.int32)
    a = q(state, u=torch.rand(n, generator=g))
    assert a.dtype == torch.int32 and int(a.min()) >= 0 and int(a.max()) <= 3
    for s in range(4):
        want = np.exp(qt[s]) / np.exp(qt[s]).sum()
        got = torch.bincount(a[state == s].long(), minlength=4).float().numpy()
        np.testing.assert_allclose(got / got.sum(), want, atol=0.012)
    assert int(q(torch.tensor([2]), u=torch.tensor([0.0]))[0]) == 0 and int(q(torch.tensor([0]), u=torch.tensor([0.999999]))[0]) == 3
    ref = qt.copy()
    for s, act, r, nx in [(0, 1, 2., 3), (3, 2, -2., 0), (0, 1, 1., 0)]:
        ref[s][act] += 0.7 * (r + 0.8 * ref[nx].max() - ref[s][act])
    q.td_update(torch.tensor([0, 3, 0]), torch.tensor([1, 2, 1]), torch.tensor([2., -2., 1.]), torch.tensor([3, 0, 0]), alpha=0.7)
    np.testing.assert_allclose(q.q_table.numpy(), ref, rtol=1e-12)


# ----------------------------------------------------------------------------------------- bench.py contract (CPU side)
@pytest.mark.parametrize('workload,kind', [('rlepso', 'port'), ('gleet', 'port'), ('rlepso', 'reference'), ('deddqn', 'reference')])
def test_bench_reference_arm_prints_one_contract_line(workload, kind):
    """`bench.py --impl reference` on the host cores, a bounded sample: ONE JSON line with the contract's keys.  kind "reference" = the
    unmodified reference staged under baseline/_ref (git-ignored; staged by __graft_entry__.build() in the build container and shipped
    with the snapshot); kind "port" = the oracle restatement."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if kind == 'reference' and not os.path.exists(os.path.join(root, 'baseline', '_ref', 'MANIFEST.json')):
        pytest.skip('baseline/_ref is not staged here')
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-gens', '2',
                        '--workload', workload, '--cpu-kind', kind], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['metric'] == 'candidate evals/sec on BBOB dim=30 pop=100' and d['value'] > 0 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == kind and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['cpu_baseline']['evals_per_s_per_core'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_staged_reference_is_byte_identical_to_its_manifest():
    """baseline/_ref (when present): every staged python source hashes to what the staging script recorded from /root/reference, and
    nothing under metabox_b200/ imports the staged reference or the oracle."""
    import hashlib
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    man = os.path.join(root, 'baseline', '_ref', 'MANIFEST.json')
    if os.path.exists(man):
        M = json.load(open(man))['python_sources_sha256']
        assert len(M) > 60
        for rel, sha in M.items():
            with open(os.path.join(root, 'baseline', '_ref', 'src', rel), 'rb') as f:
                assert hashlib.sha256(f.read()).hexdigest() == sha, rel
    for dirpath, _, files in os.walk(os.path.join(root, 'metabox_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'ref_runner' not in src and 'baseline/_ref' not in src and 'import oracle' not in src and 'from oracle' not in src, f


def test_bench_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm exits non-zero with a message instead of measuring something else."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--steps', '1', '--warmup', '1'], capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert r.returncode != 0 and not any(ln.startswith('{') for ln in r.stdout.splitlines())
    assert 'CUDA' in (r.stderr + r.stdout)


# ----------------------------------------------------------------------------------------- DEDQN engine: host logic with injected kernels
def _dedqn_cpu_engine(fids, dim, maxFEs, seed=3, noisy=True):
    """DEDQN_Engine on CPU tensors: the two CUDA kernels it composes (mbx_eval, mbx_landscape_features) are replaced by the oracle's
    functions, everything else -- the batched single-row DE arithmetic, selection, survival, accounting, curve protocol -- is the product
    code.  (Only tests can inject kernels; the product constructor needs a CUDA InstancePack.)"""
    import torch
    from metabox_b200 import dedqn as md
    from oracle import bbob as ob, landscape as ola
    tr, te = ob.make_suite('bbob', dim, 5.0, 'easy')
    insts = [[i for i in tr + te if i.fid == f][0] for f in fids]

    def ev(x, gen):
        x = x.numpy().astype(np.float64)
        return torch.as_tensor(np.stack([(ob.func_true(I, x[b]) - I.optimum).astype(np.float32) for b, I in enumerate(insts)]))

    def feats(sample, cost):
        s, c = sample.numpy(), cost.numpy()
        return torch.as_tensor(np.array([[ola.cal_fdc(s[b], c[b]), ola.cal_rie(c[b]), ola.cal_acf(c[b]), ola.cal_nop(s[b], c[b])] for b in range(len(s))]))

    eng = md.DEDQN_Engine(None, list(range(len(fids))), maxFEs, 50, seed=seed,
                          _kernels=dict(dim=dim, lb=-5., ub=5., has_opt=True, eval=ev, features=feats, noisy=noisy))
    return eng, insts


def test_dedqn_engine_host_logic_vs_oracle_with_injected_kernels():
    """Every step the engine draws its own tape on (CPU) tensors; the oracle replays exactly those draws from the engine's previous
    state: population / costs / survival / gbest bit-exact, features and reward to rounding, accounting and curve equal, lanes freeze
    when done."""
    import torch
    from oracle import bbob as ob, dedqn as od
    fids, dim, maxFEs = [1, 8, 15, 21], 10, 4400                  # 21 steps after init
    eng, insts = _dedqn_cpu_engine(fids, dim, maxFEs)
    cfg = od.Config(dim=dim, maxFEs=maxFEs)
    eng.init_population()
    t0 = {k: v.numpy() for k, v in eng.last_tape.items()}
    S = []
    for b, I in enumerate(insts):
        ev = lambda x, draws, I=I: ob.func_true(I, x)
        s, state = od.init_population(cfg, ev, I.optimum, -5., 5., od.ArrayTape({'init_u': t0['init_u'][b], 'walk_u': t0['walk_u'][b]}))
        st = eng.read_state(b)
        np.testing.assert_array_equal(st['pop'], s.pop); np.testing.assert_array_equal(st['cost'], s.cost)
        assert int(st['gbest_row']) == s.gbest_row and int(st['fes']) == s.fes == 200
        np.testing.assert_allclose(st['feature64'], state, rtol=1e-9, atol=1e-12)
        S.append(s)
    g = np.random.RandomState(0)
    for t in range(1, 24):
        action = g.randint(0, 3, size=len(fids))
        was_done = eng.done.numpy().copy()
        before = [eng.read_state(b) for b in range(len(fids))]
        feat, reward, done = eng.update(torch.as_tensor(action), check_action=True)
        tp = {k: v.numpy() for k, v in eng.last_tape.items()}
        for b, I in enumerate(insts):
            st = eng.read_state(b)
            if was_done[b]:
                for k in ('pop', 'cost', 'survival', 'fes', 'curve', 'curve_len', 'feature64'):
                    np.testing.assert_array_equal(st[k], before[b][k], err_msg=f'finished lane {b} changed {k}')
                assert float(reward[b]) == 0.0
                continue
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            rec = {'r': tp['r'][b], 'jrand': tp['jrand'][b], 'cr_u': tp['cr_u'][b], 'walk_u': tp['walk_u'][b]}
            s, state, r, d = od.update(cfg, S[b], int(action[b]), ev, I.optimum, -5., 5., od.ArrayTape(rec))
            tag = f'step {t} lane {b}'
            np.testing.assert_array_equal(st['pop'], s.pop, err_msg=tag); np.testing.assert_array_equal(st['cost'], s.cost, err_msg=tag)
            np.testing.assert_array_equal(st['survival'], s.survival, err_msg=tag)
            assert int(st['gbest_row']) == s.gbest_row and float(st['gbest_cost']) == float(s.gbest_cost), tag
            if s.gbest_row < 0:
                np.testing.assert_array_equal(st['gbest'], s.gbest, err_msg=tag)
            np.testing.assert_allclose(st['feature64'], state, rtol=1e-9, atol=1e-12, err_msg=tag)
            np.testing.assert_allclose(float(eng.reward64[b]), r, rtol=1e-13, err_msg=tag)
            assert int(st['fes']) == s.fes and int(st['log_index']) == s.log_index and bool(st['done']) == d, tag
            np.testing.assert_array_equal(st['curve'][:int(st['curve_len'])], np.array(s.curve, dtype=np.float32), err_msg=tag)
            S[b] = s
    assert bool((eng.done != 0).all()) and all(len(c) == 51 for c in eng.cost_curves())
    assert int(eng.fes.max()) == 4400
    with pytest.raises(ValueError):
        eng.update(torch.as_tensor([0, 1, 2, 3]), check_action=True)
    with pytest.raises(ValueError):
        eng.update(torch.as_tensor([0, 1]))


def test_dedqn_engine_replays_the_reference_tape():
    """Golden steps of the unmodified reference (tests/golden/dedqn.npz, one noise-free BBOB case) through the ENGINE's host logic: the
    reference's own draws are recovered by re-seeding np.random (NumpyTape) and handed to the engine as a tape."""
    import torch
    from oracle import bbob as ob, dedqn as od
    G = np.load(os.path.join(GOLDEN, 'dedqn.npz'))
    key, dim = 'f15_d10', 10
    eng, insts = _dedqn_cpu_engine([15], dim, int(G[f'{key}/maxFEs']), noisy=False)       # noise-free suite: the features reuse the stored costs
    nt = od.NumpyTape(100, dim)
    np.random.seed(int(G[f'{key}/seed0']))
    nt.init_u(); nt.walk_u()
    eng.init_population(tape={'init_u': nt.rec['init_u'][None], 'walk_u': nt.rec['walk_u'][None]})
    np.testing.assert_array_equal(eng.pop[0].numpy(), G[f'{key}/s0/pop'])
    np.testing.assert_allclose(eng.feature64[0].numpy(), G[f'{key}/s0/state'], rtol=1e-6, atol=1e-9)
    for t in range(1, 61):
        a = int(G[f'{key}/action'][t - 1])
        np.random.seed(int(G[f'{key}/seed'][t - 1]))
        j = eng.pointer
        nt = od.NumpyTape(100, dim)
        nt.r(od.N_COLS[a], j); nt.crossover(); nt.walk_u()
        tape = {'r': nt.rec['r'][None], 'jrand': np.array([nt.rec['jrand']]), 'cr_u': nt.rec['cr_u'][None], 'walk_u': nt.rec['walk_u'][None]}
        feat, reward, done = eng.update(torch.as_tensor([a]), tape=tape)
        tag = f'step {t}'
        np.testing.assert_array_equal(eng.pop[0, j].numpy(), G[f'{key}/row'][t - 1], err_msg=tag)
        np.testing.assert_allclose(float(eng.cost[0, j]), float(G[f'{key}/row_cost'][t - 1]), rtol=1e-6, err_msg=tag)
        assert float(eng.survival[0, j]) == float(G[f'{key}/row_survival'][t - 1]), tag
        np.testing.assert_allclose(eng.feature64[0].numpy(), G[f'{key}/state'][t - 1], rtol=2e-5, atol=1e-8, err_msg=tag)
        np.testing.assert_allclose(float(eng.reward64[0]), float(G[f'{key}/reward'][t - 1]), rtol=1e-12, err_msg=tag)
        assert int(eng.fes[0]) == int(G[f'{key}/fes'][t - 1]) and int(eng.log_index[0]) == int(G[f'{key}/log_index'][t - 1]), tag
        assert int(eng.curve_len[0]) == int(G[f'{key}/curve_len'][t - 1]), tag


# ----------------------------------------------------------------------------------------- RL_HPSDE engine: host logic with injected kernels
def _walk_emul():
    import ctypes as C
    import subprocess
    import tempfile
    if not hasattr(_walk_emul, 'lib'):
        d = tempfile.mkdtemp(prefix='mbx_emul_')
        so = os.path.join(d, 'walk_host.so')
        subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off', os.path.join(os.path.dirname(__file__), 'emul', 'walk_host.cpp'),
                        '-o', so], check=True)
        L = C.CDLL(so)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.pwalk_host.restype = C.c_int
        L.pwalk_host.argtypes = [dp, dp, ip, dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, dp, fp]
        _walk_emul.lib = L
    return _walk_emul.lib


def _walk_host(zone_u, r_u, rD, step_u, lb, ub):
    """tests/emul/walk_host.cpp: the device header of mbx_progressive_walk compiled for the host -> (f64 walk, f32 walk)."""
    import ctypes as C
    L = _walk_emul()
    dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    z, r, s = (np.ascontiguousarray(a, dtype=np.float64) for a in (zone_u, r_u, step_u))
    d = np.ascontiguousarray(rD, dtype=np.int32)
    B, n, D = s.shape
    out = np.empty((B, n + 1, D)); out32 = np.empty((B, n + 1, D), np.float32)
    assert L.pwalk_host(z.ctypes.data_as(dp), r.ctypes.data_as(dp), d.ctypes.data_as(ip), s.ctypes.data_as(dp), lb, ub, 10.0, B, D, n,
                        out.ctypes.data_as(dp), out32.ctypes.data_as(fp)) == 0
    return out, out32


def test_progressive_walk_device_header_on_the_host_equals_the_oracle():
    """One coordinate function for kernel and host: against the oracle's restatement of the reference walk, bit for bit, for the BBOB box
    (a step never crosses the box twice) and the protein box (narrower than a step: mirrored points stay outside, like the reference)."""
    from oracle import rl_hpsde as oh
    g = np.random.RandomState(3)
    for lb, ub, D in ((-5.0, 5.0, 10), (-5.0, 5.0, 30), (-1.5, 1.5, 12)):
        B = 5
        zone_u, r_u, rD, step_u = g.rand(B, D), g.rand(B, D), g.randint(0, D, size=B), g.rand(B, 200, D)
        out, out32 = _walk_host(zone_u, r_u, rD, step_u, lb, ub)
        for b in range(B):
            want = oh.progressive_walk(lb, ub, zone_u[b], r_u[b], int(rD[b]), step_u[b])
            np.testing.assert_array_equal(out[b], want)
            np.testing.assert_array_equal(out32[b], want.astype(np.float32))
        assert (out[:, 0][np.arange(B), rD] == np.where(zone_u[np.arange(B), rD] < 0.5, lb, ub)).all()
        assert ((out < lb) | (out > ub)).any() == (ub - lb < 10)


def _hpsde_cpu_engine(fids, dim, maxFEs, seed=3):
    """RL_HPSDE_Engine on CPU tensors: mbx_eval / mbx_landscape_features are replaced by the oracle's functions and mbx_progressive_walk
    by the host build of its own device header; everything else is the product code."""
    import torch
    from metabox_b200 import rl_hpsde as mh
    from oracle import bbob as ob, landscape as ola
    tr, te = ob.make_suite('bbob', dim, 5.0, 'easy')
    insts = [[i for i in tr + te if i.fid == f][0] for f in fids]

    def ev(x, gen):
        x = x.numpy().astype(np.float64)
        return torch.as_tensor(np.stack([(ob.func_true(I, x[b]) - I.optimum).astype(np.float32) for b, I in enumerate(insts)]))

    def feats(sample, cost, ref):
        s, c, r = sample.numpy(), cost.numpy(), ref.numpy()
        return torch.as_tensor(np.array([[ola.dfdc_r(s[b], c[b], r[b]), ola.drie_r(c[b])] for b in range(len(s))]))

    def walk(zone_u, r_u, rD, step_u):
        return torch.as_tensor(_walk_host(zone_u.numpy(), r_u.numpy(), rD.numpy(), step_u.numpy(), -5.0, 5.0)[1])

    eng = mh.RL_HPSDE_Engine(None, list(range(len(fids))), maxFEs, 50, seed=seed,
                             _kernels=dict(dim=dim, lb=-5., ub=5., has_opt=True, eval=ev, features=feats, walk=walk))
    return eng, insts


def test_rl_hpsde_engine_host_logic_vs_oracle_with_injected_kernels():
    """Every generation the engine draws its own tape on (CPU) tensors; the oracle replays exactly those draws from the engine's previous
    state: sorted population / costs bit-exact, memories and reward to rounding, population-size schedule, state, accounting and curve
    equal; a full run to the end of the budget (NP 180 -> 4)."""
    import torch
    from oracle import bbob as ob, rl_hpsde as oh
    fids, dim, maxFEs = [1, 8, 15, 21], 10, 20000
    eng, insts = _hpsde_cpu_engine(fids, dim, maxFEs)
    cfg = oh.Config(dim=dim, maxFEs=maxFEs)
    eng.init_population()
    t0 = {k: v.numpy() for k, v in eng.last_tape.items()}
    S = []
    for b, I in enumerate(insts):
        ev = lambda x, draws, I=I: ob.func_true(I, x)
        s, state = oh.init_population(cfg, ev, I.optimum, -5., 5., oh.ArrayTape({k: v[b] for k, v in t0.items()}))
        st = eng.read_state(b)
        np.testing.assert_array_equal(st['group'], s.group); np.testing.assert_array_equal(st['cost'], s.cost)
        assert int(st['feature']) == state and int(st['fes']) == s.fes == 180 + 201
        S.append(s)
    g = np.random.RandomState(0)
    gens = 0
    while not bool((eng.done != 0).all()):
        gens += 1
        action = g.randint(0, 4, size=len(fids))
        was_done = eng.done.numpy().copy()
        feat, reward, done = eng.update(torch.as_tensor(action), check_action=True)
        tp = {k: v.numpy() for k, v in eng.last_tape.items()}
        for b, I in enumerate(insts):
            if was_done[b]:
                assert float(reward[b]) == 0.0
                continue
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            s, state, r, d = oh.update(cfg, S[b], int(action[b]), ev, I.optimum, -5., 5., oh.ArrayTape({k: v[b] for k, v in tp.items()}))
            st = eng.read_state(b)
            tag = f'generation {gens} lane {b}'
            assert st['cost'].shape[0] == s.NP, tag
            np.testing.assert_array_equal(st['group'], s.group, err_msg=tag); np.testing.assert_array_equal(st['cost'], s.cost, err_msg=tag)
            np.testing.assert_allclose(st['MF'], s.MF, rtol=1e-12, err_msg=tag); np.testing.assert_allclose(st['MCr'], s.MCr, rtol=1e-12, err_msg=tag)
            assert int(st['k']) == s.k and int(st['feature']) == state and bool(st['done']) == d, tag
            np.testing.assert_allclose(float(eng.reward64[b]), r, rtol=1e-13, err_msg=tag)
            assert int(st['fes']) == s.fes and int(st['log_index']) == s.log_index, tag
            np.testing.assert_array_equal(st['curve'][:int(st['curve_len'])], np.array(s.curve, dtype=np.float32), err_msg=tag)
            S[b] = s
    assert gens >= 55 and eng.NP <= 5 and all(len(c) == 51 for c in eng.cost_curves())
    with pytest.raises(ValueError):
        eng.update(torch.as_tensor([0, 1, 2, 4]), check_action=True)


def test_rl_hpsde_distinct_index_sampler():
    """Production replacement of `generate_random_int`'s rejection loop: triples mutually distinct, != the row index, every admissible value
    equally likely (smallest population NP = 4: each row gets a permutation of the three other rows)."""
    import torch
    eng, _ = _hpsde_cpu_engine([1, 8], 10, 20000)
    for N in (4, 7, 180):
        r = torch.stack([eng._distinct(N) for _ in range(400 if N < 100 else 3)]).numpy()          # [reps, B, N, 3]
        i = np.arange(N)[None, None, :, None]
        assert r.min() >= 0 and r.max() < N and not (r == i).any()
        assert not ((r[..., 0] == r[..., 1]) | (r[..., 0] == r[..., 2]) | (r[..., 1] == r[..., 2])).any()
        if N == 7:
            for col in range(3):                                                                   # row 2 never draws 2, the other 6 values ~ 1/6 each
                cnt = np.bincount(r[:, :, 2, col].reshape(-1), minlength=N)
                assert cnt[2] == 0 and np.all(np.abs(cnt[[0, 1, 3, 4, 5, 6]] / cnt.sum() - 1 / 6) < 0.06)


# ----------------------------------------------------------------------------------------- L2L: engine + policy host logic vs golden episodes
def test_l2l_engine_and_policy_replay_reference_episodes():
    """Product host logic (L2L_Engine with the evaluation kernel replaced by the oracle's function, agent.L2L_PolicyNet with the shipped
    weights, the sweep's recurrence) through whole episodes of the unmodified reference (tests/golden/l2l.npz): costs, every-second-best
    curve, fes, return."""
    import torch
    from metabox_b200 import agent as ma, l2l as ml
    from oracle import bbob as ob
    G = np.load(os.path.join(GOLDEN, 'l2l.npz'))
    fids = [int(f) for f in G['fids'] if f < 100]
    tr, te = ob.make_suite('bbob', 10, 5.0, 'easy')
    insts = [[i for i in tr + te if i.fid == f][0] for f in fids]

    def ev(x, gen):
        x = x.numpy().astype(np.float64)
        return torch.as_tensor(np.stack([(ob.func_true(I, x[b]) - I.optimum).astype(np.float32) for b, I in enumerate(insts)]))

    net = ma.L2L_PolicyNet(10)
    net.net.load_state_dict({k[4:]: torch.as_tensor(G[k]) for k in G.files if k.startswith('net/')})
    eng = ml.L2L_Engine(None, list(range(len(fids))), _kernels=dict(dim=10, lb=-5., ub=5., has_opt=True, eval=ev))
    feat = eng.init_population()
    h, c = torch.zeros(len(fids), 10, dtype=torch.float64), torch.zeros(len(fids), 32, dtype=torch.float64)
    ys = []
    for t in range(100):
        h, c = net.step(feat, h, c)
        feat, y, done = eng.update(h)
        ys.append(y.numpy().copy())
    ys = np.array(ys)                                                  # [100, lanes]
    assert bool((done != 0).all()) and torch.all(eng.fes == 100)
    for b, f in enumerate(fids):
        np.testing.assert_allclose(ys[:, b], G[f'f{f}/y'], rtol=2e-6, err_msg=f'f{f}')
        np.testing.assert_allclose(eng.cost_curves(pad=False)[b], G[f'f{f}/cost'], rtol=2e-6)
        np.testing.assert_allclose(float(eng.episode_return()[b]), float(G[f'f{f}/ret']), rtol=1e-6)
    before = eng.fes.clone()
    eng.update(h)                                                      # finished lanes are frozen
    assert torch.equal(eng.fes, before)
    # nn.LSTM itself (the reference's call) agrees with the written-out cell
    out, (h2, c2) = net.net(feat[None, :1], (h[None, :1], c[None, :1]))
    hs, cs = net.step(feat[:1], h[:1], c[:1])
    np.testing.assert_allclose(out[0].detach().numpy(), hs.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(c2[0].detach().numpy(), cs.numpy(), rtol=1e-12, atol=1e-14)


# ----------------------------------------------------------------------------------------- NRLPSO engine: host logic with an injected evaluation kernel
def test_nrlpso_engine_host_logic_vs_oracle_with_injected_kernel():
    """The engine draws its own tape on (CPU) tensors, the oracle (pinned against the unmodified reference) replays exactly those draws
    while carrying its OWN state for 230 steps (the pointer wraps twice): rows, velocities, personal / global bests incl. the gbest view
    aliasing, neighbourhood snapshots, mutants, stagnation counters, rewards, variable evaluation counts, next-particle state."""
    import torch
    from metabox_b200 import nrlpso as mn
    from oracle import bbob as ob, nrlpso as on
    fids, dim, maxFEs = [1, 15, 21, 8], 10, 20000
    tr, te = ob.make_suite('bbob', dim, 5.0, 'easy')
    insts = [[i for i in tr + te if i.fid == f][0] for f in fids]

    def ev(x, gen):
        x = x.numpy().astype(np.float64)
        return torch.as_tensor(np.stack([(ob.func_true(I, x[b]) - I.optimum).astype(np.float32) for b, I in enumerate(insts)]))

    eng = mn.NRLPSO_Engine(None, list(range(len(fids))), maxFEs, 50, seed=5, _kernels=dict(dim=dim, lb=-5., ub=5., has_opt=True, eval=ev))
    cfg = on.Config(dim=dim, maxFEs=maxFEs)
    feat = eng.init_population()
    t0 = {k: v.numpy() for k, v in eng.last_tape.items()}
    S = []
    for b, I in enumerate(insts):
        evo = lambda x, draws, I=I: ob.func_true(I, x)
        s, state = on.init_population(cfg, evo, I.optimum, -5., 5., on.ArrayTape({k: v[b] for k, v in t0.items()}))
        assert int(feat[b]) == state
        np.testing.assert_array_equal(eng.pop[b].numpy(), s.pop); np.testing.assert_array_equal(eng.cost[b].numpy(), s.cost)
        S.append(s)
    g = np.random.RandomState(2)
    n_mut = 0
    for t in range(1, 231):
        action = g.randint(0, 4, size=len(fids))
        j = eng.pointer
        feat, reward, done = eng.update(torch.as_tensor(action), check_action=True)
        tp = {k: v.numpy() for k, v in eng.last_tape.items()}
        for b, I in enumerate(insts):
            evo = lambda x, draws, I=I: ob.func_true(I, x)
            fes0 = S[b].fes
            s, state, r, d = on.update(cfg, S[b], int(action[b]), evo, I.optimum, -5., 5., on.ArrayTape({k: v[b] for k, v in tp.items()}))
            n_mut += (s.fes - fes0 - 1) // 2
            st = eng.read_state(b)
            tag = f'step {t} lane {b}'
            for k in ('pop', 'vel', 'cost', 'pbest_pos', 'pbest_cost', 'stag', 'qstate'):
                np.testing.assert_array_equal(st[k], getattr(s, k), err_msg=f'{tag} {k}')
            assert int(st['gbest_row']) == s.gbest_row and float(st['gbest_cost']) == float(s.gbest_cost), tag
            if s.gbest_row < 0:
                np.testing.assert_array_equal(st['gbest_pos'], s.gbest_pos, err_msg=tag)
            if j == 0:
                np.testing.assert_array_equal(st['pbest_neb_index'], s.pbest_neb_index, err_msg=tag)
                np.testing.assert_array_equal(st['gbest_neb_index'], s.gbest_neb_index, err_msg=tag)
                np.testing.assert_array_equal(st['pbest_neb'], s.pbest_neb, err_msg=tag)
                assert abs(float(st['w']) - s.w) <= 1e-15 and abs(float(st['r_w']) - s.r_w) <= 1e-15, tag
            assert int(feat[b]) == state and float(reward[b]) == r and bool(done[b]) == d, tag
            assert int(st['fes']) == s.fes and int(st['log_index']) == s.log_index, tag
            np.testing.assert_array_equal(st['curve'][:int(st['curve_len'])], np.array(s.curve, dtype=np.float32), err_msg=tag)
            S[b] = s
    assert n_mut >= 5 and eng.pointer == 30
    with pytest.raises(ValueError):
        eng.update(torch.as_tensor([0, 1, 2, 4]), check_action=True)


# ----------------------------------------------------------------------------------------- LES engine: host logic with an injected evaluation kernel
def test_les_engine_host_logic_vs_oracle_and_golden_generations():
    """(1) Golden generations of the unmodified reference (tests/golden/les.npz) through the ENGINE's batched update rule: the reference's
    normal draws are recovered by re-seeding np.random and handed over as a tape.  (2) Free-running lanes: every generation the oracle
    replays the engine's own draws from the engine's previous state; a whole episode ends with the curve protocol intact."""
    import torch
    from metabox_b200 import les as ml
    from oracle import bbob as ob, les as ol
    G = np.load(os.path.join(GOLDEN, 'les.npz'))
    tr, te = ob.make_suite('bbob', 10, 5.0, 'easy')
    insts = [[i for i in tr + te if i.fid == f][0] for f in (1, 15)]

    def ev(x, gen):
        x = x.numpy().astype(np.float64)
        return torch.as_tensor(np.stack([ob.func_true(I, x[b]) for b, I in enumerate(insts)]))

    mk = lambda maxFEs: ml.LES_Engine(None, [0, 1], maxFEs, 50, seed=3, _kernels=dict(dim=10, lb=-5., ub=5., has_opt=True, eval=ev)).set_params(G['attn'], G['mlp'])
    # (1) replay of the reference
    eng = mk(20000)
    keys = ['f1_d10', 'f15_d10']
    mu_u, zn = [], []
    for key in keys:
        np.random.seed(int(G[f'{key}/seed0']))
        mu_u.append(np.random.rand(10)); zn.append(np.random.standard_normal((16, 10)))
    eng.init_population(tape={'mu_u': np.stack(mu_u), 'normal': np.stack(zn)})
    for b, key in enumerate(keys):
        np.testing.assert_allclose(eng.parents[b].numpy(), G[f'{key}/s0/parents'], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(eng.cost[b].numpy(), G[f'{key}/s0/cost'], rtol=1e-12)
    for t in range(1, 41):
        zn = []
        for key in keys:
            np.random.seed(int(G[f'{key}/seed'][t - 1])); zn.append(np.random.standard_normal((16, 10)))
        eng.update(tape={'normal': np.stack(zn)})
        for b, key in enumerate(keys):
            for k in ('parents', 'mu', 'sigma', 'Pc', 'Ps', 'cost'):
                np.testing.assert_allclose(getattr(eng, k)[b].numpy(), G[f'{key}/{k}'][t - 1], rtol=2e-5, atol=1e-7, err_msg=f'{key} generation {t} {k}')
            assert int(eng.fes[b]) == int(G[f'{key}/fes'][t - 1]) and int(eng.log_index[b]) == int(G[f'{key}/log_index'][t - 1])
            np.testing.assert_allclose(float(eng.gbest[b]), float(G[f'{key}/gbest'][t - 1]), rtol=1e-6)
    # (2) free-running, oracle from the engine's previous state
    eng = mk(800)
    eng.init_population()
    for t in range(1, 50):
        prev = [eng.read_state(b) for b in range(2)]
        was_done = eng.done.numpy().copy()
        t_before = eng.t
        eng.update()
        zn = eng.last_tape['normal'].numpy()
        for b, I in enumerate(insts):
            if was_done[b]:
                continue
            q = prev[b]
            S = ol.State(parents=q['parents'], cost=q['cost'], t=t_before, gbest=float(q['gbest']), Pc=q['Pc'], Ps=q['Ps'], mu=q['mu'], sigma=q['sigma'],
                         fes=int(q['fes']), log_index=int(q['log_index']), curve=list(q['curve'][:int(q['curve_len'])]))
            N = ol.generation(S, G['attn'], G['mlp'], 10, 800, 16, lambda x, d, I=I: ob.func_true(I, x), I.optimum, -5., 5., normal=zn[b])
            st = eng.read_state(b)
            for k in ('parents', 'mu', 'sigma', 'Pc', 'Ps', 'cost'):
                np.testing.assert_allclose(st[k], getattr(N, k), rtol=2e-5, atol=1e-7, err_msg=f'generation {t} lane {b} {k}')
            assert int(st['fes']) == N.fes and int(st['log_index']) == N.log_index and bool(st['done']) == N.done
    assert bool((eng.done != 0).all()) and int(eng.fes.max()) == 800 and all(len(c) == 51 for c in eng.cost_curves())
    r = eng.episode_return().numpy()
    assert np.all(r >= 0) and np.all(r < 1)
    eng2 = mk(800)
    eng2.init_population()
    gb, ret, done = eng2.run_episode()
    assert bool((done != 0).all()) and int(eng2.fes.max()) == 800
    with pytest.raises(AssertionError):
        eng2.set_params(G['attn'][:-1], G['mlp'])


def test_fused_policy_headers_with_random_weights():
    """The host builds of the four policy headers against their torch modules for freshly initialised (random) weights and inputs: the packing
    order and every layer are exercised with values that do not come from one trained checkpoint."""
    import ctypes as C
    import torch
    from metabox_b200 import agent as ma
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
    for seed in (1, 2, 3):
        torch.manual_seed(seed)
        g = np.random.RandomState(seed)
        # GLEET attention actor
        actor = ma.GLEET_Actor()
        with torch.no_grad():
            for p in actor.parameters():
                p.mul_(0.7).sub_(0.2)                                     # the reference initialises the attention matrices in [0, 1): mix the signs
        NP = int(g.choice([3, 50, 100]))
        x = g.randn(2, NP, 27).astype(np.float32)
        act, mu, sg, lg = _emul_forward(actor.pack().numpy(), x)
        m2, s2, l2 = actor.distribution(torch.as_tensor(x))
        np.testing.assert_allclose(lg, l2.detach().numpy(), rtol=5e-4, atol=5e-5)
        np.testing.assert_allclose(mu, m2.detach().numpy(), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(sg, s2.detach().numpy(), rtol=2e-4, atol=2e-5)
        # DE-DDQN Q-network
        net = ma.DE_DDQN_Net()
        L = _emul_lib('deddqn_policy')
        w = np.ascontiguousarray(net.pack().numpy())
        xs = np.ascontiguousarray(g.randn(13, 99).astype(np.float32))
        a = np.empty(13, np.int32); q = np.empty((13, 4), np.float32)
        assert L.ddq_host_forward(xs.ctypes.data_as(fp), w.ctypes.data_as(fp), a.ctypes.data_as(ip), q.ctypes.data_as(fp), 13) == 0
        want = net(torch.as_tensor(xs)).detach().numpy()
        np.testing.assert_allclose(q, want, rtol=1e-4, atol=1e-5)
        np.testing.assert_array_equal(a, np.argmax(want, axis=1))
        # RLEPSO actor
        ra = ma.RLEPSO_Actor()
        L = _emul_lib('rlepso_policy')
        w = np.ascontiguousarray(ra.pack().numpy())
        xs = np.ascontiguousarray(g.rand(9).astype(np.float32)); nz = np.ascontiguousarray(g.randn(9, 35).astype(np.float32))
        act = np.empty((9, 35), np.float32)
        assert L.rlp_host_forward(xs.ctypes.data_as(fp), w.ctypes.data_as(fp), nz.ctypes.data_as(fp), act.ctypes.data_as(fp), None, None, 9) == 0
        np.testing.assert_allclose(act, ra(torch.as_tensor(xs)[:, None], noise=torch.as_tensor(nz)).numpy(), rtol=1e-5, atol=2e-6)
        # RL_PSO policy network
        dim = int(g.choice([10, 12, 30]))
        rp = ma.RL_PSO_PolicyNet(dim)
        L = _emul_lib('pso_policies')
        w = np.ascontiguousarray(rp.pack().numpy())
        xs = np.ascontiguousarray(g.randn(7, 2 * dim).astype(np.float32)); nz = np.ascontiguousarray((2 * g.randn(7)).astype(np.float32))
        act = np.empty(7, np.float32)
        assert L.rpp_host_forward(xs.ctypes.data_as(fp), w.ctypes.data_as(fp), nz.ctypes.data_as(fp), act.ctypes.data_as(fp), None, None, 7, 2 * dim) == 0
        np.testing.assert_allclose(act, rp(torch.as_tensor(xs), noise=torch.as_tensor(nz)[:, None]).numpy()[:, 0], rtol=1e-5, atol=2e-6)


def test_tester_checks_policy_classes_against_the_backbone():
    """A policy object of another backbone is rejected up front (it used to fail deep inside torch with a dtype error); unknown callables
    (user policies) pass."""
    from metabox_b200 import agent as ma, tester as mt
    with pytest.raises(ValueError):
        mt._check_policy('nrlpso', ma.RLEPSO_Actor())
    with pytest.raises(ValueError):
        mt._check_policy('rlepso', ma.DE_DDQN_Net())
    mt._check_policy('rlepso', ma.RLEPSO_Actor())
    mt._check_policy('qlpso', ma.QLPSO_QTable(np.zeros((4, 4))))
    mt._check_policy('nrlpso', ma.QLPSO_QTable(np.zeros((4, 4))))
    mt._check_policy('gleet', lambda state: state[..., 0])


def test_gather_results_rejects_more_rows_than_promised():
    """`n_max` is the caller's promise about the largest per-rank row count: breaking it is an error, not a truncated gather."""
    import torch
    import torch.distributed as dist
    from metabox_b200 import sweep
    res = dict(curve=torch.zeros(5, 51), fes=torch.zeros(5, dtype=torch.int32), ret=torch.zeros(5))
    assert sweep.gather_results(res, 1, n_max=3).shape == (5, 53)          # world 1: no collective, nothing to check
    if not (dist.is_available() and dist.is_initialized()):
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29577')
        dist.init_process_group('gloo', rank=0, world_size=1)
        try:
            with pytest.raises(ValueError):
                sweep.gather_results(res, 2, n_max=3)
        finally:
            dist.destroy_process_group()
