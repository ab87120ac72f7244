"""GPU: the BASELINE.json configurations at their full shapes, through size-independent properties plus oracle replays of
sampled lanes (C2 RLEPSO bbob D10 x 64 runs; C3 DE-DDQN bbob-difficult D30 x 256 runs; C4 LDE bbob-noisy D30 NP100 sharded
over 4 ranks; C5 RLEPSO on 280 protein instances x 51 runs)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _weights(group):
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    return {k[len(group) + 1:]: G[k] for k in G.files if k.startswith(group + '/')}


def _rows(res):
    from metabox_b200 import sweep
    return sweep.gather_results(res, 1).cpu().numpy()


def _curve_ok(r, maxFEs, has_opt=True):
    assert np.all(np.isfinite(r[:, :51])) and np.all(np.diff(r[:, :51], axis=1) <= 0), 'gbest curves must be non-increasing'
    if has_opt:
        assert np.all((r[:, 51] >= maxFEs) | (r[:, 50] <= 1e-8))
    else:
        assert np.all(r[:, 51] >= maxFEs)


def test_c2_rlepso_bbob_d10_64_runs_full_rollouts():
    """24 functions x 64 runs x full 20000-FE rollouts with the shipped policy; striped == properties hold; 3 sampled lanes
    re-run alone (same seed, same global lane id, element-wise policy) reproduce the batched rows bit for bit."""
    import torch
    from metabox_b200 import agent as ma, problem as mp, sweep
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    probs = tr.data + te.data
    pack = mp.InstancePack(probs)
    inst, _ = sweep.lane_table(len(probs), 64)
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), _weights('rlepso')).cuda().eval()
    torch.manual_seed(0)
    r = _rows(sweep.StripedSweep(lambda ln, lb: sweep.RLEPSO_Sweep(pack, ln, actor, 20000, 50, seed=3, lane_base=lb), inst, 0, 3).run())
    assert r.shape == (1536, 53)
    _curve_ok(r, 20000)
    assert np.all(r[:, 51] <= 20000 + 200) and np.all(np.abs(r[:, 52]) <= 201)
    solved = (r[:, 50] <= 1e-8).reshape(24, 64).mean(1)
    k1 = [p.fid for p in probs].index(1)
    assert solved[k1] == 1.0 and solved.sum() >= 2, solved                     # Sphere always converges with the shipped policy

    # placement independence of the ENGINE: an element-wise policy (cuBLAS picks different GEMM kernels for different batch
    # sizes, so an MLP's float32 actions are not bit-identical between a 1-lane and a 1536-lane batch)
    class Elementwise:
        def __call__(self, x):
            k = torch.arange(1, 36, device=x.device, dtype=torch.float32)
            return torch.clamp(0.5 + 0.45 * torch.sin(x * k * 7.0), 0, 1)
    full = _rows(sweep.RLEPSO_Sweep(pack, inst, Elementwise(), 20000, 50, seed=3).run())
    for lane in (5, 700, 1535):
        one = _rows(sweep.RLEPSO_Sweep(pack, inst[lane:lane + 1], Elementwise(), 20000, 50, seed=3, lane_base=lane).run())
        np.testing.assert_array_equal(one[0], full[lane])


def test_c3_deddqn_difficult_d30_256_runs_lockstep():
    """18 functions x 256 runs = 4608 lanes in lockstep for 300 steps with the shipped Q-network."""
    import torch
    from metabox_b200 import agent as ma, optimizer as mo, problem as mp
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 30, 5, difficulty='difficult')
    assert len(te.data) == 18
    pack = mp.InstancePack(te.data)
    eng = mo.DE_DDQN_Engine(pack, np.repeat(np.arange(18), 256), 60000, 50, seed=8)
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), _weights('deddqn')).cuda().eval()
    feat = eng.init_population()
    prev_cost, prev_g = eng.cost.clone(), eng.c3[:, 0].clone()
    for step in range(300):
        a = torch.argmax(net(torch.nan_to_num(feat)), dim=-1).to(torch.int32)
        feat, reward, done = eng.update(a)
        if step % 50 == 49:
            assert torch.all(eng.cost <= prev_cost)                           # `<=` replacement never worsens a row
            assert torch.equal(eng.c3[:, 0], torch.minimum(prev_g, eng.cost.min(dim=1).values))
            assert torch.all(eng.fes == 100 + step + 1) and torch.all(reward >= 0) and torch.all(torch.isfinite(feat))
            prev_cost, prev_g = eng.cost.clone(), eng.c3[:, 0].clone()
    assert int(eng.ints[:, 7].sum()) == 0 and torch.all(eng.X.abs() <= 5)


def test_c4_lde_noisy_np100_shards_equal_the_whole():
    """30 noisy functions x 8 runs, NP = 100: the 4 contiguous shards a 4-rank job would run (shard_range) give exactly the
    rows of the unsharded sweep (Philox keyed by the global lane id); element-wise policy."""
    import torch
    from metabox_b200 import agent as ma, problem as mp, sweep
    tr, te = mp.BBOB_Dataset.get_datasets('bbob-noisy', 30, 5)
    probs = tr.data + te.data
    assert len(probs) == 30
    pack = mp.InstancePack(probs)
    class Mean:                      # element-wise (batch-size invariant) stand-in for the LSTM policy
        NP, cell = 100, 50

        def step(self, x, h, c):
            f = x[:, :100]
            return torch.clip(torch.cat([0.5 + 0.4 * torch.sin(9.0 * f), 0.6 + 0.4 * torch.cos(5.0 * f)], dim=1), 0, 1), h, c
    inst, _ = sweep.lane_table(30, 8)
    with torch.no_grad():
        whole = _rows(sweep.LDE_Sweep(pack, inst, Mean(), 3000, 50, seed=6, NP=100).run())
        parts = []
        for rank in range(4):
            lo, hi = sweep.shard_range(len(inst), rank, 4)
            parts.append(_rows(sweep.LDE_Sweep(pack, inst[lo:hi], Mean(), 3000, 50, seed=6, lane_base=lo, NP=100).run()))
    np.testing.assert_array_equal(np.concatenate(parts), whole)
    _curve_ok(whole, 3000)
    assert np.all(whole[:, 51] == 3000)


def test_c5_rlepso_protein_280_instances_51_runs():
    """280 (synthetic, shipped-shape) protein instances x 51 runs = 14280 lanes, maxFEs 1000 (10 generations, 5 log points);
    two sampled lanes are replayed generation by generation by the oracle (Philox twin): positions bit-exact."""
    import torch
    from metabox_b200 import agent as ma, optimizer as mo, problem as mp, sweep
    from oracle import protein as op, rlepso as orl
    probs = [mp.synthetic_protein(1000 + k) for k in range(280)]
    pack = mp.InstancePack(probs)
    inst, _ = sweep.lane_table(280, 51)
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), _weights('rlepso')).cuda().eval()
    torch.manual_seed(0)
    r = _rows(sweep.StripedSweep(lambda ln, lb: sweep.RLEPSO_Sweep(pack, ln, actor, 1000, 5, seed=4, lane_base=lb), inst, 0, 3).run())
    assert r.shape == (14280, 53)
    _curve_ok(r, 1000, has_opt=False)
    assert np.all(r[:, 51] >= 1000) and np.all(r[:, 51] <= 1000 + 10 * 100)
    assert np.all(r[:, 5] == r[:, 50])                                        # 6 samples, then padded with the last one
    # oracle replay of lanes 3 and 14000
    seed = 4
    eng = mo.RLEPSO_Engine(pack, inst, 1000, 5, seed=seed)
    eng.init_population()
    cfg = orl.Config(dim=12, maxFEs=1000, n_logpoint=5)
    g = np.random.RandomState(0)
    for gen in range(3):
        action = g.rand(len(inst), 35).astype(np.float32)
        prev = {b: eng.read_state(b) for b in (3, 14000)}
        eng.update(torch.as_tensor(action).to(eng.device))
        for b in (3, 14000):
            P = probs[inst[b]]
            I = op.ProteinInstance(P.problem_id, P.coor_init, P.q_vec, P.e_vec, P.r_vec, P.basis, P.eigval)
            ev = lambda x, draws, I=I: op.energy(I, x)
            p = prev[b]
            S = orl.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], pni=p['pni'],
                          gbest_x=p['gbest_x'], gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']), no_improve=0,
                          fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, _, _, _ = orl.update(cfg, S, action[b], ev, None, op.LB, op.UB, orl.PhiloxTape(cfg.NP, 12, seed, b, gen))
            st = eng.read_state(b)
            np.testing.assert_array_equal(st['x'], N.x); np.testing.assert_array_equal(st['v'], N.v)
            np.testing.assert_array_equal(st['pbest_x'], N.pbest_x)
            np.testing.assert_allclose(st['c_cost'], N.c_cost, rtol=1e-6, atol=1e-9)
            assert int(st['fes']) == N.fes and int(st['gbest_idx']) == N.gbest_idx
