"""GPU: batched rollout drivers (metabox_b200/sweep.py) -- CUDA-graph replay == eager launches, recurrent policy
state is reset between sweeps, and complete rollouts with the reference's shipped policies land where the
reference's own runs land (SURVEY.md 8c anchors / for_review/AEI_data summary statistics)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _weights(group):
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    return {k[len(group) + 1:]: G[k] for k in G.files if k.startswith(group + '/')}


def _pack(dim=10, suit='bbob'):
    from metabox_b200 import problem as mp
    tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
    return mp.InstancePack(tr.data + te.data), tr.data + te.data


def _rows(sw):
    from metabox_b200 import sweep
    return sweep.gather_results(sw.run(), 1).cpu().numpy()


def test_deddqn_sweep_graph_equals_eager():
    import torch
    from metabox_b200 import agent as ma, sweep
    pack, probs = _pack()
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), _weights('deddqn')).cuda().eval()
    inst, _ = sweep.lane_table(len(probs), 2)
    a = sweep.DE_DDQN_Sweep(pack, inst, net, 460, 50, seed=5, graph=True, steps_per_graph=7, poll=70)
    b = sweep.DE_DDQN_Sweep(pack, inst, net, 460, 50, seed=5, graph=False, poll=50)
    ra, rb, ra2 = _rows(a), _rows(b), _rows(a)
    np.testing.assert_array_equal(ra, rb)              # deterministic policy + counter-based RNG: bit-identical
    np.testing.assert_array_equal(ra, ra2)             # replaying the captured graph from a fresh init
    assert np.all(ra[:, 51] == 460) and np.all(np.diff(ra[:, :51], axis=1) <= 0)
    assert a.launches >= 360 and int(a.engine.ints[:, 7].sum()) == 0


class _MeanActor:
    """RLEPSO actor without sampling noise (deterministic -> graph and eager runs are comparable bit for bit)."""

    def __init__(self, actor):
        self.actor = actor

    def __call__(self, x):
        import torch
        mu, _ = self.actor.distribution(x)
        return torch.clamp(mu, 0, 1)


def test_rlepso_sweep_graph_equals_eager():
    from metabox_b200 import agent as ma, sweep
    pack, probs = _pack()
    actor = _MeanActor(ma.load_reference_weights(ma.RLEPSO_Actor(), _weights('rlepso')).cuda().eval())
    inst, _ = sweep.lane_table(len(probs), 2)
    a = sweep.RLEPSO_Sweep(pack, inst, actor, 3000, 50, seed=9, graph=True, poll=8)
    b = sweep.RLEPSO_Sweep(pack, inst, actor, 3000, 50, seed=9, graph=False, poll=8)
    ra, rb = _rows(a), _rows(b)
    np.testing.assert_array_equal(ra, rb)
    assert np.all((ra[:, 51] >= 3000) | (ra[:, 50] <= 1e-8)) and np.all(ra[:, 51] < 3100 + 100)
    assert np.all(np.abs(ra[:, 52]) <= 31)             # return = sum of +-1 over <= 30 generations


def test_striped_sweep_on_two_streams_equals_the_single_sweep():
    """Lane stripes on separate CUDA streams / graphs: Philox streams are keyed by the global lane id, so the rows are
    bit-identical to the unstriped sweep (element-wise policy)."""
    from metabox_b200 import agent as ma, sweep
    import torch
    pack, probs = _pack()

    def actor(x):                    # element-wise => bit-identical whatever the batch size (an MLP's GEMM kernel choice is not)
        k = torch.arange(1, 36, device=x.device, dtype=torch.float32)
        return torch.clamp(0.5 + 0.45 * torch.sin(x * k * 7.0), 0, 1)
    inst, _ = sweep.lane_table(len(probs), 3)
    one = sweep.RLEPSO_Sweep(pack, inst, actor, 3000, 50, seed=9, lane_base=7, graph=True, poll=8)
    two = sweep.StripedSweep(lambda ln, lb: sweep.RLEPSO_Sweep(pack, ln, actor, 3000, 50, seed=9, lane_base=lb, graph=True, poll=8),
                             inst, lane_base=7, n_stripes=3)
    r1, r2, r2b = _rows(one), _rows(two), _rows(two)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(r2, r2b)
    assert two.launches >= 3 * 30


class _MeanLDE:
    def __init__(self, net):
        self.net, self.NP, self.cell = net, net.NP, net.cell

    def step(self, x, h, c):
        import torch
        mu, sigma, h2, c2 = self.net.lstm_cell(x, h, c)
        return torch.clip(mu, 0, 1), h2, c2


def test_lde_sweep_graph_equals_eager_and_resets_lstm_state():
    import torch
    from metabox_b200 import agent as ma, sweep
    pack, probs = _pack(suit='bbob-noisy')
    with torch.no_grad():
        net = _MeanLDE(ma.load_reference_weights(ma.LDE_PolicyNet(), _weights('lde')).cuda().eval())
        inst, _ = sweep.lane_table(len(probs), 2)
        a = sweep.LDE_Sweep(pack, inst, net, 2000, 50, seed=4, graph=True, poll=8)
        b = sweep.LDE_Sweep(pack, inst, net, 2000, 50, seed=4, graph=False, poll=8)
        ra, rb, ra2 = _rows(a), _rows(b), _rows(a)
    np.testing.assert_array_equal(ra, rb)
    np.testing.assert_array_equal(ra, ra2)             # (h, c) zeroed at reset like lde_agent.py:150-151
    assert np.all(ra[:, 51] == 2000)


def test_full_rollouts_land_where_the_reference_lands():
    """Sphere (F1) dim 10 with the shipped bbob_easy policies.  Reference anchors (seed 0): RLEPSO
    fes 8500 / 8.3e-9, LDE fes 9250 / 9.4e-9, DE_DDQN fes 9859 / 6.1e-9; published 51-run means: RLEPSO 7102 FEs."""
    import torch
    from metabox_b200 import agent as ma, problem as mp, sweep
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    sphere = [p for p in tr.data + te.data if p.fid == 1]
    pack = mp.InstancePack(sphere)
    runs = 48
    inst, _ = sweep.lane_table(1, runs)
    torch.manual_seed(0)
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), _weights('rlepso')).cuda().eval()
    r = _rows(sweep.RLEPSO_Sweep(pack, inst, actor, 20000, 50, seed=1))
    assert np.all(r[:, 50] <= 1e-8) and 5500 <= r[:, 51].mean() <= 11000, ('rlepso', r[:, 51].mean(), r[:, 50].max())
    lde = ma.load_reference_weights(ma.LDE_PolicyNet(), _weights('lde')).cuda().eval()
    r = _rows(sweep.LDE_Sweep(pack, inst, lde, 20000, 50, seed=2))
    assert np.all(r[:, 50] <= 1e-8) and 6500 <= r[:, 51].mean() <= 11500, ('lde', r[:, 51].mean(), r[:, 50].max())
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), _weights('deddqn')).cuda().eval()
    r = _rows(sweep.DE_DDQN_Sweep(pack, inst[:16], net, 20000, 50, seed=3))
    assert np.median(r[:, 50]) <= 1e-6 and 7000 <= r[:, 51].mean() <= 20000, ('deddqn', r[:, 51].mean(), np.median(r[:, 50]))
