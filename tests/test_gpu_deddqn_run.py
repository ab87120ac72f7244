"""GPU: the persistent DE-DDQN kernel (`mbx_deddqn_run`, SURVEY.md K6: n_steps function evaluations per launch, lane state resident in shared
memory, the Q-network inside the kernel) against the validated per-step path (`mbx_deddqn_policy` ; `mbx_deddqn_step`, itself pinned
against the oracle / the reference in tests/test_gpu_deddqn.py, tests/test_gpu_benchmark_shapes.py and tests/test_gpu_z_policies.py).

The two paths share their device code (csrc/mbx_deddqn.cuh, mbx_deddqn_policy.cuh), so the comparison is BIT FOR BIT on every state array,
the features, the cost curves and the accumulated return -- across launch boundaries, generation roll-overs, lanes that finish inside a
launch, lanes of different functions (and noise models) sharing a CTA, lane counts that do not fill the last CTA, and the protein energy."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
KEYS = ('X', 'cost', 'X_gbest', 'X_prebest', 'c3', 'ints', 'r', 'n_tot', 'n_succ', 'om_sum', 'om_max', 'win', 'fes', 'log_index', 'curve_len',
        'curve', 'done', 'feature')


def _net():
    import torch
    from metabox_b200 import agent as ma
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    W = {k[len('deddqn/'):]: G[k] for k in G.files if k.startswith('deddqn/')}
    return ma.DE_DDQN_FusedNet(ma.load_reference_weights(ma.DE_DDQN_Net(), W).cuda().eval())


def _problems(cases, dim):
    from metabox_b200 import problem as mp
    out = []
    for suit, f in cases:
        tr, te = mp.BBOB_Dataset.get_datasets(suit, dim, 5)
        out.append([p for p in tr.data + te.data if p.fid == f][0])
    return out


def _compare(a, b, tag):
    for k in KEYS:
        x, y = getattr(a, k).cpu().numpy(), getattr(b, k).cpu().numpy()
        if k == 'feature':                      # NaN features (degenerate lanes) must be NaN on both sides
            np.testing.assert_array_equal(np.isnan(x), np.isnan(y), err_msg=f'{tag} {k} NaN pattern')
            x, y = np.nan_to_num(x), np.nan_to_num(y)
        np.testing.assert_array_equal(x, y, err_msg=f'{tag} {k}')


def _both(pack, lanes, maxFEs, seed, chunks, NP=None, lane_base=3):
    """per-step path for sum(chunks) steps vs the persistent kernel launched once per chunk."""
    import torch
    from metabox_b200 import optimizer as mo
    net = _net()
    A = mo.DE_DDQN_Engine(pack, lanes, maxFEs, 50, seed=seed, lane_base=lane_base, NP=NP)
    Bn = mo.DE_DDQN_Engine(pack, lanes, maxFEs, 50, seed=seed, lane_base=lane_base, NP=NP)
    A.init_population(); Bn.init_population()
    _compare(A, Bn, 'init')
    ra = torch.zeros(A.B, device='cuda'); rb = torch.zeros(A.B, device='cuda')
    done_steps = 0
    for n in chunks:
        for _ in range(n):
            _, rew, _ = A.update(net(A.feature))
            ra += torch.nan_to_num(rew, nan=0.0, posinf=0.0, neginf=0.0)
        Bn.run_steps(net.packed, rb, n)
        done_steps += n
        _compare(A, Bn, f'after {done_steps} steps')
        np.testing.assert_array_equal(ra.cpu().numpy(), rb.cpu().numpy(), err_msg=f'return after {done_steps} steps')
    return A, Bn


@pytest.mark.parametrize('dim', [10, 20, 30])
def test_persistent_kernel_equals_policy_plus_step_bit_for_bit(dim):
    from metabox_b200 import problem as mp
    cases = [('bbob', 1), ('bbob', 7), ('bbob', 16), ('bbob', 21), ('bbob-noisy', 104), ('bbob-noisy', 118), ('bbob-noisy', 129),
             ('bbob', 23), ('bbob', 20), ('bbob', 5)]
    probs = _problems(cases, dim)
    pack = mp.InstancePack(probs)
    lanes = list(range(len(probs))) + [0, 3, 6]             # 13 lanes: the last CTA of the persistent kernel is ragged, CTAs mix functions
    A, Bn = _both(pack, lanes, 2000 * dim, 0xD0D0 + dim, chunks=[1, 7, 95, 130])      # crosses the pointer roll-over at 100 and 200
    assert int(A.ints[:, 1].min()) >= 2                     # every lane went through two generation boundaries


def test_lanes_that_finish_inside_a_launch_freeze_exactly_like_the_step_path():
    """maxFEs = 260: every lane ends after 160 steps, inside the second launch; Sphere lanes at dim 10 may stop earlier on 1e-8 -- either
    way the persistent kernel must stop the lane at the same evaluation, write the same final curve point and leave it untouched after."""
    from metabox_b200 import problem as mp
    probs = _problems([('bbob', 1), ('bbob', 15), ('bbob-noisy', 107), ('bbob', 8), ('bbob', 24)], 10)
    pack = mp.InstancePack(probs)
    A, Bn = _both(pack, list(range(5)), 260, 77, chunks=[100, 100, 50])
    assert bool((A.done != 0).all()) and int(A.fes.max()) == 260


def test_other_population_sizes_and_the_protein_energy():
    from metabox_b200 import problem as mp
    from test_gpu_rlepso import _problem
    probs = _problems([('bbob', 3), ('bbob', 12), ('bbob', 22)], 10)
    pack = mp.InstancePack(probs)
    _both(pack, [0, 1, 2, 1, 0], 20000, 5, chunks=[60, 45], NP=37)
    G = np.load(os.path.join(GOLDEN, 'protein.npz'))
    names = sorted({k.split('_', 2)[2] for k in G.files if k.startswith('coor_init_')})[:2]
    prot = [_problem(n, 12)[0] for n in names]
    ppack = mp.InstancePack(prot)
    _both(ppack, [0, 1, 1], 1000, 9, chunks=[40, 75])


def test_persistent_sweep_equals_the_graph_sweep_and_is_the_default():
    from metabox_b200 import problem as mp, sweep
    probs = _problems([('bbob', 1), ('bbob', 6), ('bbob', 17), ('bbob-noisy', 121)], 10)
    pack = mp.InstancePack(probs)
    inst, _ = sweep.lane_table(len(probs), 3)
    net = _net()
    s1 = sweep.DE_DDQN_Sweep(pack, inst, net, 3000, 50, seed=12)
    assert s1.persistent
    r1 = sweep.gather_results(s1.run(), 1).cpu().numpy()
    r2 = sweep.gather_results(sweep.DE_DDQN_Sweep(pack, inst, net, 3000, 50, seed=12, persistent=False).run(), 1).cpu().numpy()
    np.testing.assert_array_equal(r1, r2)
    assert np.all((r1[:, 51] == 3000) | (r1[:, 50] <= 1e-8)) and s1.launches < 20
    r3 = sweep.gather_results(s1.run(), 1).cpu().numpy()    # a reused sweep starts from a fresh population
    np.testing.assert_array_equal(r1, r3)
    with pytest.raises(ValueError):
        sweep.DE_DDQN_Sweep(pack, inst, net.net, 3000, 50, seed=12, persistent=True)
