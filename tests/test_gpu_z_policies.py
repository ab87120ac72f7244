"""GPU: the GLEET attention actor and the RL_PSO policy network (metabox_b200/agent.py) on the device -- known-answer forwards of the
unmodified reference (tests/golden/agents2.npz), agreement with the oracle's head-by-head restatement on states the fused GLEET step
itself produced, and complete sweeps driven by these policies (CUDA-graph replay == eager launches).

The torch-module tests ran green on a B200 (profiles/r1i_gpu_policies_tests.log).  The tests of the FUSED policy kernels (second half of
this file) were written when round 1's GPU minutes were almost spent; its last seconds went into the fused GLEET kernel: the known-answer
test passed on a B200 and so did every edge-size / error check of the following test up to a bug in the TEST (it indexed a Dropout module;
fixed since) -- profiles/r1i_gpu_fused_gleet_policy_tests.log.  Still without a device run: the fused DE-DDQN Q-network kernel, the two
sweeps on fused policies and GLEET_Trainer; their arithmetic is verified on the host (tests/test_host_cpu.py compiles the same device
headers with g++).  The file sorts last so that `pytest -x` reaches every GPU test that has already been validated on a B200 first."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def _weights(group):
    G = np.load(os.path.join(GOLDEN, 'agents2.npz'))
    return {k[len(group) + 1:]: G[k] for k in G.files if k.startswith(group + '/')}, G


def _pack(dim=10):
    from metabox_b200 import problem as mp
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    return mp.InstancePack(tr.data + te.data), tr.data + te.data


class _Mean:
    """A policy without sampling noise (graph and eager runs become comparable bit for bit)."""

    def __init__(self, net, pick):
        self.net, self.pick = net, pick

    def __call__(self, x):
        import torch
        return torch.clamp(self.pick(self.net.distribution(x)), 0, 1)


def test_gleet_actor_and_rl_pso_net_known_answers_on_the_device():
    import torch
    from metabox_b200 import agent as ma
    for grp in ('gleet', 'gleet0'):
        W, G = _weights(grp)
        actor = ma.load_reference_weights(ma.GLEET_Actor(), W).cuda().eval()
        for name in ('a', 'b', 'r'):
            x = torch.as_tensor(G[f'{grp}_kat/{name}/x']).cuda()
            with torch.no_grad():
                mu, sigma, lg = actor.distribution(x)
            np.testing.assert_allclose(lg.cpu().numpy(), G[f'{grp}_kat/{name}/logits'], rtol=5e-4, atol=5e-5)
            np.testing.assert_allclose(mu.cpu().numpy(), G[f'{grp}_kat/{name}/mu'][..., 0], rtol=2e-4, atol=2e-5)
            np.testing.assert_allclose(sigma.cpu().numpy(), G[f'{grp}_kat/{name}/sigma'][..., 0], rtol=2e-4, atol=2e-5)
            a = actor(x)
            assert a.shape == x.shape[:2] and float(a.min()) >= 0 and float(a.max()) <= 1
    W2, _ = _weights('rl_pso')
    net = ma.load_reference_weights(ma.RL_PSO_PolicyNet(10), W2).cuda().eval()
    a = net(torch.as_tensor(G['rl_pso_kat/x']).cuda(), noise=torch.as_tensor(G['rl_pso_kat/eps']).cuda())
    np.testing.assert_allclose(a.cpu().numpy(), G['rl_pso_kat/action'], rtol=1e-5, atol=1e-6)


def test_gleet_actor_on_states_of_the_fused_step_vs_oracle_restatement():
    """States come out of the CUDA GLEET step (Philox mode); the device actor's (mu, sigma) must agree with the oracle's head-by-head
    float32 CPU restatement on exactly those states, generation after generation."""
    import torch
    from metabox_b200 import agent as ma, optimizer as mo
    from oracle import agents as oag
    W, G = _weights('gleet')
    actor = ma.load_reference_weights(ma.GLEET_Actor(), W).cuda().eval()
    w_cpu = oag.load_weights(G, 'gleet')
    pack, probs = _pack()
    eng = mo.GLEET_Engine(pack, np.arange(len(probs)), 20000, 50, seed=11)
    feat = eng.init_population()
    for gen in range(6):
        with torch.no_grad():
            mu, sigma, _ = actor.distribution(feat)
        omu, osg, _ = oag.gleet_distribution(w_cpu, feat.cpu().numpy())
        np.testing.assert_allclose(mu.cpu().numpy(), omu.numpy(), rtol=5e-4, atol=5e-5, err_msg=f'gen {gen}')
        np.testing.assert_allclose(sigma.cpu().numpy(), osg.numpy(), rtol=5e-4, atol=5e-5, err_msg=f'gen {gen}')
        feat, _, _ = eng.update(torch.clamp(mu, 0, 1))


def test_gleet_sweep_with_the_attention_actor():
    from metabox_b200 import agent as ma, sweep
    W, _ = _weights('gleet')
    actor = ma.load_reference_weights(ma.GLEET_Actor(), W).cuda().eval()
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    rows = lambda sw: sweep.gather_results(sw.run(), 1).cpu().numpy()
    mean = _Mean(actor, lambda d: d[0])
    ra = rows(sweep.GLEET_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=True, poll=8))
    rb = rows(sweep.GLEET_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=False, poll=8))
    np.testing.assert_array_equal(ra, rb)
    # sampled actions (noise from the sweep's own generator, captured with the graph): complete, monotone cost curves
    sw = sweep.GLEET_Sweep(pack, inst, actor, 3000, 50, seed=9, graph=True, poll=8)
    r1, r2 = rows(sw), rows(sw)
    for r in (r1, r2):
        assert np.all((r[:, 51] >= 3000) | (r[:, 50] <= 1e-8)) and np.all(r[:, 51] <= 3100)
        assert np.all(np.diff(r[:, :51], axis=1) <= 0) and np.isfinite(r).all()
    assert not np.array_equal(r1[:, :51], r2[:, :51])            # a second sweep draws fresh action noise
    # striped over two streams: same Philox keys per global lane; the deterministic policy gives identical rows
    st = sweep.StripedSweep(lambda ln, lb: sweep.GLEET_Sweep(pack, ln, mean, 3000, 50, seed=9, lane_base=lb, poll=8), inst, 0, 2)
    np.testing.assert_array_equal(rows(st), ra)


def test_rl_pso_sweep_with_the_reference_policy_net():
    from metabox_b200 import agent as ma, sweep
    W, _ = _weights('rl_pso')
    net = ma.load_reference_weights(ma.RL_PSO_PolicyNet(10), W).cuda().eval()
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    rows = lambda sw: sweep.gather_results(sw.run(), 1).cpu().numpy()
    mean = _Mean(net, lambda d: d[0].reshape(-1))
    ra = rows(sweep.RL_PSO_Sweep(pack, inst, mean, 460, 50, seed=5, graph=True, steps_per_graph=7, poll=70))
    rb = rows(sweep.RL_PSO_Sweep(pack, inst, mean, 460, 50, seed=5, graph=False, poll=50))
    np.testing.assert_array_equal(ra, rb)
    r = rows(sweep.RL_PSO_Sweep(pack, inst, net, 460, 50, seed=5, graph=True, steps_per_graph=7, poll=70))
    assert np.all((r[:, 51] == 460) | (r[:, 50] <= 1e-8)) and np.all(np.diff(r[:, :51], axis=1) <= 0) and np.isfinite(r).all()


# ------------------------------------------------------------------------------------------------ fused policy kernel (mbx_gleet_policy)
def test_fused_gleet_policy_kernel_known_answers_and_torch_module():
    """One CTA per swarm: against forwards of the unmodified reference (both checkpoints; random, ragged and reference-produced states) and
    against the torch module on the device under the same action noise."""
    import torch
    from metabox_b200 import agent as ma
    g = torch.Generator(device='cpu').manual_seed(4)
    for grp in ('gleet', 'gleet0'):
        W, G = _weights(grp)
        actor = ma.load_reference_weights(ma.GLEET_Actor(), W).cuda().eval()
        fused = ma.GLEET_FusedActor(actor)
        for name in ('a', 'b', 'r'):
            x = torch.as_tensor(G[f'{grp}_kat/{name}/x']).cuda()
            noise = torch.randn(x.shape[:2], generator=g).cuda()
            act, mu, sigma, lg = fused(x, noise=noise, outputs=True)
            np.testing.assert_allclose(lg.cpu().numpy(), G[f'{grp}_kat/{name}/logits'], rtol=2e-4, atol=3e-5)
            np.testing.assert_allclose(mu.cpu().numpy(), G[f'{grp}_kat/{name}/mu'][..., 0], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(sigma.cpu().numpy(), G[f'{grp}_kat/{name}/sigma'][..., 0], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(act.cpu().numpy(), actor(x, noise=noise).cpu().numpy(), rtol=2e-4, atol=3e-5)
            assert torch.equal(fused(x, noise=noise), act)                                 # outputs are optional, the action is the same
            assert torch.equal(fused(x), torch.clamp(mu, 0, 1))                            # null noise -> the mean action


def test_fused_gleet_policy_kernel_edge_sizes_errors_and_refresh():
    import torch
    from metabox_b200 import _lib, agent as ma
    W, _ = _weights('gleet0')
    actor = ma.load_reference_weights(ma.GLEET_Actor(), W).cuda().eval()
    fused = ma.GLEET_FusedActor(actor)
    g = torch.Generator(device='cpu').manual_seed(8)
    for B, NP in ((1, 1), (3, 2), (2, 5), (5, 37), (2, 128), (700, 100)):
        x = torch.rand(B, NP, 27, generator=g).cuda()
        act, mu, sigma, lg = fused(x, outputs=True)
        with torch.no_grad():
            m2, s2, l2 = actor.distribution(x)
        np.testing.assert_allclose(lg.cpu().numpy(), l2.cpu().numpy(), rtol=5e-4, atol=5e-5, err_msg=f'B {B} NP {NP}')
        np.testing.assert_allclose(mu.cpu().numpy(), m2.cpu().numpy(), rtol=2e-4, atol=2e-5, err_msg=f'B {B} NP {NP}')
        np.testing.assert_allclose(sigma.cpu().numpy(), s2.cpu().numpy(), rtol=2e-4, atol=2e-5, err_msg=f'B {B} NP {NP}')
    assert fused(torch.zeros(0, 100, 27).cuda()).shape == (0, 100)                         # empty batch: nothing to launch
    with pytest.raises(_lib.MbxError):
        fused(torch.zeros(1, 129, 27).cuda())
    with pytest.raises(ValueError):
        fused(torch.zeros(1, 100, 26).cuda())
    with pytest.raises(ValueError):
        fused(torch.zeros(1, 100, 27, dtype=torch.float64).cuda())
    # the kernel reads the packed buffer on every launch: refresh() after a parameter update changes the next action
    x = torch.rand(4, 100, 27, generator=g).cuda()
    before = fused(x).clone()
    with torch.no_grad():
        [m for m in actor.mu_net.net if isinstance(m, torch.nn.Linear)][-1].bias.add_(0.5)
    assert torch.equal(fused(x), before)
    after = fused.refresh()(x)
    assert not torch.equal(after, before)
    np.testing.assert_allclose(after.cpu().numpy(), actor(x, noise=torch.zeros(4, 100).cuda()).cpu().numpy(), rtol=2e-4, atol=3e-5)


def test_gleet_sweep_with_the_fused_policy_kernel():
    """Complete GLEET rollouts with fused step + fused policy: graph replay == eager launches (mean actions), sampled sweeps are sane."""
    from metabox_b200 import agent as ma, sweep
    W, _ = _weights('gleet')
    fused = ma.GLEET_FusedActor(ma.load_reference_weights(ma.GLEET_Actor(), W).cuda().eval())
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    rows = lambda sw: sweep.gather_results(sw.run(), 1).cpu().numpy()
    mean = lambda st: fused(st)                                                          # no noise -> deterministic
    ra = rows(sweep.GLEET_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=True, poll=8))
    rb = rows(sweep.GLEET_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=False, poll=8))
    np.testing.assert_array_equal(ra, rb)
    r = rows(sweep.GLEET_Sweep(pack, inst, fused, 3000, 50, seed=9, graph=True, poll=8))
    assert np.all((r[:, 51] >= 3000) | (r[:, 50] <= 1e-8)) and np.all(r[:, 51] <= 3100)
    assert np.all(np.diff(r[:, :51], axis=1) <= 0) and np.isfinite(r).all()


# ------------------------------------------------------------------------------------------------ fused Q-network kernel (mbx_deddqn_policy)
def test_fused_deddqn_policy_kernel_vs_reference_q_values_and_torch_net():
    import torch
    from metabox_b200 import agent as ma
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), {k[len('deddqn/'):]: G[k] for k in G.files if k.startswith('deddqn/')}).cuda().eval()
    fused = ma.DE_DDQN_FusedNet(net)
    act, q = fused(torch.as_tensor(G['deddqn_kat/x']).cuda(), q=True)
    np.testing.assert_allclose(q.cpu().numpy(), G['deddqn_kat/q'], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(act.cpu().numpy(), np.argmax(G['deddqn_kat/q'], axis=1))
    g = torch.Generator(device='cpu').manual_seed(3)
    for B in (1, 7, 8, 9, 771, 4608):
        x = (torch.rand(B, 99, generator=g) * 2 - 0.5)
        if B > 8:
            x[3, 17] = float('nan'); x[5, 0] = float('inf'); x[B - 1, 98] = float('-inf')
        x = x.cuda()
        act, q = fused(x, q=True)
        with torch.no_grad():
            want = net(torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0))
        np.testing.assert_allclose(q.cpu().numpy(), want.cpu().numpy(), rtol=2e-4, atol=2e-5, err_msg=f'B {B}')
        top2 = torch.topk(want, 2, dim=1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4                                      # lanes without a near-tie between the two best actions
        assert torch.equal(act[clear].long(), torch.argmax(want, dim=1)[clear]) and act.dtype == torch.int32
        assert torch.equal(fused(x), act)
    assert fused(torch.zeros(0, 99).cuda()).shape == (0,)
    with pytest.raises(ValueError):
        fused(torch.zeros(4, 98).cuda())


def test_deddqn_sweep_with_the_fused_q_network():
    """Complete DE-DDQN rollouts, fused step + fused Q-network: graph replay == eager launches, and the same rows as the sweep driven by the
    torch network wherever no lane ever met a near-tie (the two Q implementations differ in summation order only)."""
    from metabox_b200 import agent as ma, sweep
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    net = ma.load_reference_weights(ma.DE_DDQN_Net(), {k[len('deddqn/'):]: G[k] for k in G.files if k.startswith('deddqn/')}).cuda().eval()
    fused = ma.DE_DDQN_FusedNet(net)
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    rows = lambda sw: sweep.gather_results(sw.run(), 1).cpu().numpy()
    ra = rows(sweep.DE_DDQN_Sweep(pack, inst, fused, 460, 50, seed=5, graph=True, steps_per_graph=7, poll=70))
    rb = rows(sweep.DE_DDQN_Sweep(pack, inst, fused, 460, 50, seed=5, graph=False, poll=50))
    np.testing.assert_array_equal(ra, rb)
    assert np.all(ra[:, 51] == 460) and np.all(np.diff(ra[:, :51], axis=1) <= 0) and np.isfinite(ra).all()
    rt = rows(sweep.DE_DDQN_Sweep(pack, inst, net, 460, 50, seed=5, graph=True, steps_per_graph=7, poll=70))
    same = np.all(ra == rt, axis=1)
    assert same.mean() >= 0.5, f'only {same.sum()} of {len(same)} lanes agree with the torch-driven sweep'


# ------------------------------------------------------------------------------------------------ GLEET PPO training on the lane machinery
def test_gleet_trainer_collects_on_the_gpu_and_updates():
    import argparse
    import torch
    from metabox_b200 import agent as ma, problem as mp, trainer as mt
    from metabox_b200.optimizer import GLEET_Engine
    cfg = argparse.Namespace(dim=10, maxFEs=2500, n_logpoint=50, log_interval=50)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    W, _ = _weights('gleet0')
    Wc, _ = _weights('gleet0_critic')
    actor, critic = ma.load_reference_weights(ma.GLEET_Actor(), W), ma.load_reference_weights(ma.GLEET_Critic(), Wc)
    T = mt.GLEET_Trainer(cfg, tr.data, actor=actor, critic=critic, seed=5)
    p0 = [p.detach().clone() for p in T._params()]
    info = T.train_episodes([0, 1, 2, 3])
    assert info['segments'] == 3 and info['learn_steps'] == 9 and not info['done']          # 25 generations -> 3 segments x 3 epochs
    assert len(info['gbest']) == 4 and all(g <= n for g, n in zip(info['gbest'], info['normalizer']))
    assert all(np.isfinite(r) and r >= 0 for r in info['return']) and np.isfinite(info['reinforce_loss']) and np.isfinite(info['baseline_loss'])
    moved = max(float((a - b.detach()).abs().max()) for a, b in zip(p0, T._params()))
    assert 1e-6 < moved < 1e-2 and all(torch.isfinite(p).all() for p in T._params())       # Adam, lr 1e-4, 9 clipped steps
    eng = GLEET_Engine(T.pack, np.arange(3, dtype=np.int32), 2500, 50, 50, seed=1)
    eng.init_population()
    seg = T.collect_segment(eng)
    assert seg['states'].shape == (10, 3, 100, 27) and seg['actions'].shape == (10, 3, 100) and seg['logp'].shape == (10, 3)
    assert torch.all(seg['weight'] == 1) and torch.all(seg['rewards'] >= 0) and seg['next_state'].shape == (3, 100, 27)
    assert torch.all((seg['actions'] >= 0) & (seg['actions'] <= 1))


# ------------------------------------------------------------------------------------------------ DEDQN backbone (composed step, metabox_b200/dedqn.py)
def test_dedqn_engine_on_the_device_vs_oracle():
    """The composed DEDQN step with the real kernels (mbx_eval, mbx_landscape_features): every step the oracle replays the engine's own
    draws from the engine's previous state -- replaced rows bit-exact, costs to float32 rounding, landscape features of the engine's own
    (walk, cost) pair against the oracle's functions."""
    import torch
    from metabox_b200 import dedqn as md, problem as mp
    from oracle import bbob as ob, dedqn as od, landscape as ola
    dim, maxFEs, fids = 10, 20000, [1, 3, 9, 17, 21, 24]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    pack = mp.InstancePack([probs[f] for f in fids])
    eng = md.DEDQN_Engine(pack, list(range(len(fids))), maxFEs, 50, seed=12)
    cfg = od.Config(dim=dim, maxFEs=maxFEs)
    eng.init_population()
    assert eng.feature.shape == (len(fids), 4) and torch.all(eng.fes == 200)
    g = np.random.RandomState(1)
    for t in range(1, 31):
        action = g.randint(0, 3, size=len(fids))
        prev = [eng.read_state(b) for b in range(len(fids))]
        p = eng.pointer
        feat, reward, done = eng.update(torch.as_tensor(action).cuda(), check_action=True)
        tp = {k: v.cpu().numpy() for k, v in eng.last_tape.items()}
        smp, cst = eng.last_sample.cpu().numpy(), eng.last_cost.cpu().numpy()
        for b, f in enumerate(fids):
            I = oin[f]
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            q = prev[b]
            S = od.State(pop=q['pop'], cost=q['cost'], survival=q['survival'], gbest_row=int(q['gbest_row']), gbest=q['gbest'],
                         gbest_cost=np.float32(q['gbest_cost']), pointer=p, fes=int(q['fes']), log_index=int(q['log_index']),
                         curve=list(q['curve'][:int(q['curve_len'])]))
            rec = {'r': tp['r'][b], 'jrand': tp['jrand'][b], 'cr_u': tp['cr_u'][b], 'walk_u': tp['walk_u'][b]}
            N, state, r, d = od.update(cfg, S, int(action[b]), ev, I.optimum, -5., 5., od.ArrayTape(rec))
            st = eng.read_state(b)
            tag = f'step {t} lane {b} (fid {f})'
            np.testing.assert_array_equal(st['pop'], N.pop, err_msg=tag)
            np.testing.assert_allclose(st['cost'], N.cost, rtol=1e-6, atol=1e-10, err_msg=tag)
            np.testing.assert_array_equal(st['survival'], N.survival, err_msg=tag)
            assert int(st['gbest_row']) == N.gbest_row and int(st['fes']) == N.fes and int(st['log_index']) == N.log_index, tag
            np.testing.assert_allclose(cst[b], N.cost.astype(np.float64), rtol=1e-6, atol=1e-10, err_msg=tag)      # the re-evaluated population
            own = np.array([ola.cal_fdc(smp[b], cst[b]), ola.cal_rie(cst[b]), ola.cal_acf(cst[b]), ola.cal_nop(smp[b], cst[b])])
            np.testing.assert_allclose(st['feature64'], own, rtol=1e-9, atol=1e-12, err_msg=tag)
            np.testing.assert_allclose(st['feature64'], state, rtol=1e-3, atol=1e-6, err_msg=tag)
            np.testing.assert_allclose(float(eng.reward64[b]), r, rtol=1e-12, err_msg=tag)
    with pytest.raises(ValueError):
        eng.update(torch.as_tensor([0, 1, 2, 3, 0, 1]).cuda(), check_action=True)


def test_dedqn_sweep_tester_and_dropin():
    import argparse
    import torch
    from metabox_b200 import agent as ma, environment as me, optimizer as mo, problem as mp, sweep
    W, _ = _weights('dedqn')
    net = ma.load_reference_weights(ma.DEDQN_Net(), W).cuda().eval()
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    sw = sweep.DEDQN_Sweep(pack, inst, net, 3000, 50, seed=4)
    r1 = sweep.gather_results(sw.run(), 1).cpu().numpy()
    r2 = sweep.gather_results(sweep.DEDQN_Sweep(pack, inst, net, 3000, 50, seed=4).run(), 1).cpu().numpy()
    np.testing.assert_array_equal(r1, r2)                                                  # same seed -> same rows
    assert np.all((r1[:, 51] == 3000) | (r1[:, 50] <= 1e-8)) and np.all(np.diff(r1[:, :51], axis=1) <= 0) and np.isfinite(r1).all()
    assert np.all(r1[:, 52] > 0) and sw.launches >= 2 * 14
    # drop-in: PBO_Env + DEDQN_Optimizer with the reference's signatures
    cfg = argparse.Namespace(dim=10, maxFEs=3000, n_logpoint=50, log_interval=60)
    env = me.PBO_Env(probs[0], mo.DEDQN_Optimizer(cfg))
    np.random.seed(0)
    state = env.reset()
    assert state.shape == (4,) and state.dtype == np.float64 and env.optimizer.fes == 200 and len(env.optimizer.cost) == 1
    done, n = False, 0
    while not done:
        state, reward, done = env.step(int(np.random.randint(0, 3)))
        assert isinstance(reward, float) and 0 < reward <= 1 and state.shape == (4,)
        n += 1
    assert n == 14 and env.optimizer.fes == 3000 and len(env.optimizer.cost) == 16           # ONE log point per step (:183-185) + the final one
    with pytest.raises(ValueError):
        env.step(3)


# ------------------------------------------------------------------------------------------------ fused RLEPSO actor kernel (mbx_rlepso_policy)
def test_fused_rlepso_policy_kernel_and_sweep():
    import torch
    from metabox_b200 import agent as ma, sweep
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), {k[len('rlepso/'):]: G[k] for k in G.files if k.startswith('rlepso/')}).cuda().eval()
    fused = ma.RLEPSO_FusedActor(actor)
    act, mu, sigma = fused(torch.as_tensor(G['rlepso_kat/x']).cuda(), outputs=True)
    np.testing.assert_allclose(mu.cpu().numpy(), G['rlepso_kat/mu'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sigma.cpu().numpy(), G['rlepso_kat/sigma'], rtol=1e-5, atol=1e-6)
    g = torch.Generator(device='cpu').manual_seed(2)
    for B in (1, 3, 4, 5, 6144):                                                           # ragged last CTA (4 lanes per CTA)
        x = torch.rand(B, 1, generator=g).cuda(); nz = torch.randn(B, 35, generator=g).cuda()
        np.testing.assert_allclose(fused(x, noise=nz).cpu().numpy(), actor(x, noise=nz).cpu().numpy(), rtol=2e-5, atol=3e-6, err_msg=f'B {B}')
        assert torch.equal(fused(x.reshape(-1)), torch.clamp(fused(x, outputs=True)[1], 0, 1))
    assert fused(torch.zeros(0, 1).cuda()).shape == (0, 35)
    with pytest.raises(ValueError):
        fused(torch.zeros(4, 2).cuda())
    # complete sweeps: graph replay == eager launches with the mean action, sampled sweep sane
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    rows = lambda sw: sweep.gather_results(sw.run(), 1).cpu().numpy()
    mean = lambda st: fused(st)
    ra = rows(sweep.RLEPSO_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=True, poll=8))
    rb = rows(sweep.RLEPSO_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=False, poll=8))
    np.testing.assert_array_equal(ra, rb)
    r = rows(sweep.RLEPSO_Sweep(pack, inst, fused, 3000, 50, seed=9, graph=True, poll=8))
    assert np.all((r[:, 51] >= 3000) | (r[:, 50] <= 1e-8)) and np.all(r[:, 51] < 3200) and np.all(np.abs(r[:, 52]) <= 31)


# ------------------------------------------------------------------------------------------------ fused RL_PSO / QLPSO policy kernels
def test_fused_rl_pso_and_qlpso_policy_kernels_and_sweeps():
    import torch
    from metabox_b200 import agent as ma, sweep
    W, G = _weights('rl_pso')
    net = ma.load_reference_weights(ma.RL_PSO_PolicyNet(10), W).cuda().eval()
    fused = ma.RL_PSO_FusedPolicy(net)
    act, mu, sigma = fused(torch.as_tensor(G['rl_pso_kat/x']).cuda(), noise=torch.as_tensor(G['rl_pso_kat/eps']).cuda(), outputs=True)
    np.testing.assert_allclose(mu.cpu().numpy(), G['rl_pso_kat/mu'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sigma.cpu().numpy(), G['rl_pso_kat/sigma'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(act.cpu().numpy(), G['rl_pso_kat/action'], rtol=1e-5, atol=2e-6)
    g = torch.Generator(device='cpu').manual_seed(3)
    for dim, B in ((10, 1), (12, 5), (30, 6144)):
        torch.manual_seed(dim)
        n2 = ma.RL_PSO_PolicyNet(dim).cuda().eval()
        f2 = ma.RL_PSO_FusedPolicy(n2)
        x = (torch.rand(B, 2 * dim, generator=g) * 2 - 0.5).cuda(); nz = (3 * torch.randn(B, 1, generator=g)).cuda()
        np.testing.assert_allclose(f2(x, noise=nz).cpu().numpy(), n2(x, noise=nz).cpu().numpy(), rtol=2e-5, atol=3e-6, err_msg=f'dim {dim}')
    qt = ma.QLPSO_QTable(np.array([[0., 1., 2., 3.], [0., 0., 0., 0.], [5., 0., 0., 0.], [-1., -1., 4., -1.]])).cuda()
    fq = ma.QLPSO_FusedQTable(qt)
    state = torch.randint(0, 4, (5000,), generator=g, dtype=torch.int32).cuda(); u = torch.rand(5000, generator=g).cuda()
    assert torch.equal(fq(state, u=u), qt(state, u=u)) and fq(state).dtype == torch.int32
    # complete sweeps with both policies fused (one library kernel for the policy + one for the step per environment step)
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    rows = lambda sw: sweep.gather_results(sw.run(), 1).cpu().numpy()
    for make in (lambda graph: sweep.RL_PSO_Sweep(pack, inst, lambda st: fused(st), 460, 50, seed=5, graph=graph, steps_per_graph=7, poll=70),
                 lambda graph: sweep.QLPSO_Sweep(pack, inst, lambda st: fq(st, u=torch.full((st.shape[0],), 0.37, device=st.device)), 460, 50,
                                                 seed=5, graph=graph, steps_per_graph=7, poll=70)):
        ra, rb = rows(make(True)), rows(make(False))
        np.testing.assert_array_equal(ra, rb)                                              # deterministic policies: graph == eager
        assert np.all((ra[:, 51] == 460) | (ra[:, 50] <= 1e-8)) and np.all(np.diff(ra[:, :51], axis=1) <= 0) and np.isfinite(ra).all()
    r = rows(sweep.RL_PSO_Sweep(pack, inst, fused, 460, 50, seed=5, graph=True, steps_per_graph=7, poll=70))
    assert np.isfinite(r).all() and np.all(np.diff(r[:, :51], axis=1) <= 0)
    r = rows(sweep.QLPSO_Sweep(pack, inst, fq, 460, 50, seed=5, graph=True, steps_per_graph=7, poll=70))
    assert np.isfinite(r).all() and np.all(np.diff(r[:, :51], axis=1) <= 0)


def test_tester_with_fused_policies(tmp_path):
    """`config.fused_policies = True`: the Tester swaps every policy for its single-kernel twin and writes the reference's layouts."""
    import argparse
    import torch
    from metabox_b200 import agent as ma, tester as mt
    G1 = np.load(os.path.join(GOLDEN, 'agents.npz'))
    sub = lambda G, pre: {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    W2, G2 = _weights('gleet')
    pol = {'RLEPSO_Agent': ma.load_reference_weights(ma.RLEPSO_Actor(), sub(G1, 'rlepso/')).cuda().eval(),
           'DE_DDQN_Agent': ma.load_reference_weights(ma.DE_DDQN_Net(), sub(G1, 'deddqn/')).cuda().eval(),
           'GLEET_Agent': ma.load_reference_weights(ma.GLEET_Actor(), W2).cuda().eval(),
           'RL_PSO_Agent': ma.load_reference_weights(ma.RL_PSO_PolicyNet(10), sub(G2, 'rl_pso/')).cuda().eval(),
           'QLPSO_Agent': ma.QLPSO_QTable(np.array([[1., 0, 0, 2], [0, 1, 0, 0], [0, 0, 1, 0], [2, 0, 0, 1]])).cuda(),
           'DEDQN_Agent': ma.load_reference_weights(ma.DEDQN_Net(), sub(G2, 'dedqn/')).cuda().eval()}
    cfg = argparse.Namespace(problem='bbob', difficulty='easy', dim=10, maxFEs=600, n_logpoint=50, log_interval=12, upperbound=5,
                             test_log_dir=str(tmp_path / 'test') + '/', agent_for_cp=list(pol), t_optimizer_for_cp=[], fused_policies=True)
    T = mt.Tester(cfg, pol)
    assert [type(T.policies[k]).__name__ for k in pol] == ['RLEPSO_FusedActor', 'DE_DDQN_FusedNet', 'GLEET_FusedActor', 'RL_PSO_FusedPolicy',
                                                          'QLPSO_FusedQTable', 'DEDQN_Net']
    R = T.test(n_runs=2, random_search_baseline=False)['test']
    for p in R['cost']:
        assert set(R['cost'][p]) == set(pol)
        for a in R['cost'][p]:
            assert len(R['cost'][p][a]) == 2 and all(len(c) == 51 for c in R['cost'][p][a])
            assert all(f >= 600 or c[-1] <= 1e-8 for f, c in zip(R['fes'][p][a], R['cost'][p][a]))


def test_population_sweeps_with_several_steps_per_graph():
    """Graphs of 4 generations: finished lanes idle through surplus generations, the rows equal the one-generation-per-graph sweep."""
    from metabox_b200 import agent as ma, sweep
    G = np.load(os.path.join(GOLDEN, 'agents.npz'))
    actor = ma.load_reference_weights(ma.RLEPSO_Actor(), {k[len('rlepso/'):]: G[k] for k in G.files if k.startswith('rlepso/')}).cuda().eval()
    mean = _Mean(actor, lambda d: d[0])
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    rows = lambda sw: sweep.gather_results(sw.run(), 1).cpu().numpy()
    r1 = rows(sweep.RLEPSO_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=True, poll=8))
    r4 = rows(sweep.RLEPSO_Sweep(pack, inst, mean, 3000, 50, seed=9, graph=True, poll=8, steps_per_graph=4))
    np.testing.assert_array_equal(r1, r4)
    W, _ = _weights('gleet')
    gm = _Mean(ma.load_reference_weights(ma.GLEET_Actor(), W).cuda().eval(), lambda d: d[0])
    g1 = rows(sweep.GLEET_Sweep(pack, inst, gm, 3000, 50, seed=9, graph=True, poll=8))
    g3 = rows(sweep.GLEET_Sweep(pack, inst, gm, 3000, 50, seed=9, graph=True, poll=9, steps_per_graph=3))
    np.testing.assert_array_equal(g1, g3)


# ------------------------------------------------------------------------------------------------ RL_HPSDE backbone (composed step, metabox_b200/rl_hpsde.py)
def test_rl_hpsde_engine_on_the_device_vs_oracle():
    """The composed RL_HPSDE generation with the real kernels (mbx_eval, mbx_progressive_walk, mbx_landscape_features): every generation the
    oracle replays the engine's own draws from the engine's previous state."""
    import torch
    from metabox_b200 import problem as mp, rl_hpsde as mh
    from oracle import bbob as ob, landscape as ola, rl_hpsde as oh
    dim, maxFEs, fids = 10, 20000, [1, 3, 9, 17, 21, 24]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    pack = mp.InstancePack([probs[f] for f in fids])
    eng = mh.RL_HPSDE_Engine(pack, list(range(len(fids))), maxFEs, 50, seed=12)
    cfg = oh.Config(dim=dim, maxFEs=maxFEs)
    eng.init_population()
    assert eng.feature.shape == (len(fids),) and torch.all(eng.fes == 180 + 201) and eng.NP == 180
    # the walk kernel against the oracle's restatement, bit for bit
    tp = {k: v.cpu().numpy() for k, v in eng.last_tape.items()}
    for b in range(len(fids)):
        want = oh.progressive_walk(-5., 5., tp['zone_u'][b], tp['r_u'][b], int(tp['rD'][b]), tp['step_u'][b])
        np.testing.assert_array_equal(eng.last_sample[b].cpu().numpy(), want.astype(np.float32))
    g = np.random.RandomState(1)
    for t in range(1, 13):
        action = g.randint(0, 4, size=len(fids))
        prev = [eng.read_state(b) for b in range(len(fids))]
        feat, reward, done = eng.update(torch.as_tensor(action).cuda(), check_action=True)
        tp = {k: v.cpu().numpy() for k, v in eng.last_tape.items()}
        smp, cst = eng.last_sample.cpu().numpy().astype(np.float64), eng.last_cost.cpu().numpy()
        for b, f in enumerate(fids):
            I = oin[f]
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            q = prev[b]
            S = oh.State(group=q['group'], cost=q['cost'], MF=q['MF'], MCr=q['MCr'], k=int(q['k']), gbest=np.float32(q['gbest']), fes=int(q['fes']),
                         log_index=int(q['log_index']), curve=list(q['curve'][:int(q['curve_len'])]))
            N, state, r, d = oh.update(cfg, S, int(action[b]), ev, I.optimum, -5., 5., oh.ArrayTape({k: v[b] for k, v in tp.items()}))
            st = eng.read_state(b)
            tag = f'generation {t} lane {b} (fid {f})'
            assert st['cost'].shape[0] == N.NP, tag
            np.testing.assert_allclose(st['cost'], N.cost, rtol=1e-6, atol=1e-10, err_msg=tag)
            if np.all(np.diff(N.cost) > 1e-5 * np.abs(N.cost[1:])):                            # no near-ties: the sorted order is unambiguous
                np.testing.assert_array_equal(st['group'], N.group, err_msg=tag)
            np.testing.assert_allclose(st['MF'], N.MF, rtol=1e-4, atol=1e-7, err_msg=tag)
            np.testing.assert_allclose(st['MCr'], N.MCr, rtol=1e-4, atol=1e-7, err_msg=tag)
            assert int(st['k']) == N.k and int(st['fes']) == N.fes and int(st['log_index']) == N.log_index, tag
            own = (ola.dfdc_r(smp[b], cst[b], st['group'][0].astype(np.float64)), ola.drie_r(cst[b]))
            np.testing.assert_allclose([float(st['dfdc']), float(st['drie'])], own, rtol=1e-9, atol=1e-12, err_msg=tag)
            np.testing.assert_allclose(float(eng.reward64[b]), r, atol=2.0 / N.NP, err_msg=tag)   # a last-bit cost difference may flip one strict comparison
    assert eng.NP < 180


def test_rl_hpsde_sweep_tester_and_dropin():
    import argparse
    import torch
    from metabox_b200 import agent as ma, environment as me, optimizer as mo, sweep
    qt = ma.RL_HPSDE_QTable(_weights('rl_hpsde')[1]['rl_hpsde_q']).cuda()                  # the shipped agent's Q-table
    pack, probs = _pack()
    inst, _ = sweep.lane_table(len(probs), 2)
    r1 = sweep.gather_results(sweep.RL_HPSDE_Sweep(pack, inst, qt, 4000, 50, seed=4).run(), 1).cpu().numpy()
    r2 = sweep.gather_results(sweep.RL_HPSDE_Sweep(pack, inst, ma.fuse(qt), 4000, 50, seed=4).run(), 1).cpu().numpy()
    for r in (r1, r2):
        assert np.all((r[:, 51] >= 4000) | (r[:, 50] <= 1e-8)) and np.all(np.diff(r[:, :51], axis=1) <= 0) and np.isfinite(r).all()
        assert np.all(r[:, 52] > 0)
    np.testing.assert_array_equal(r1, r2)                       # same uniforms, same table: the fused sampler picks the same actions
    cfg = argparse.Namespace(dim=10, maxFEs=4000, n_logpoint=50, log_interval=80)
    env = me.PBO_Env(probs[0], mo.RL_HPSDE_Optimizer(cfg))
    np.random.seed(0)
    state = env.reset()
    assert state in (0, 1, 2, 3) and env.optimizer.fes == 180 + 201
    done, n = False, 0
    while not done:
        state, reward, done = env.step(int(np.random.randint(0, 4)))
        assert isinstance(reward, float) and 0 <= reward <= 1 and state in (0, 1, 2, 3)
        n += 1
    assert 9 <= n <= 16 and env.optimizer.fes >= 4000      # every generation charges NP + 201 evaluations, NP shrinking from 180 towards 4
    with pytest.raises(ValueError):
        env.step(4)


# ------------------------------------------------------------------------------------------------ L2L (metabox_b200/l2l.py)
def test_l2l_sweep_and_dropin_replay_reference_episodes():
    """Whole episodes of the unmodified reference L2L agent (tests/golden/l2l.npz, shipped weights) through the sweep on the device and through
    the drop-in optimizer: every-second-best curve, fes and return."""
    import argparse
    import torch
    from metabox_b200 import agent as ma, environment as me, optimizer as mo, problem as mp, sweep
    G = np.load(os.path.join(GOLDEN, 'l2l.npz'))
    fids = [int(f) for f in G['fids'] if f < 100]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    pack = mp.InstancePack([probs[f] for f in fids])
    net = ma.L2L_PolicyNet(10)
    net.net.load_state_dict({k[4:]: torch.as_tensor(G[k]) for k in G.files if k.startswith('net/')})
    net = net.cuda().eval()
    rows = sweep.gather_results(sweep.L2L_Sweep(pack, np.arange(len(fids), dtype=np.int32), net).run(), 1).cpu().numpy()
    for b, f in enumerate(fids):
        want = list(G[f'f{f}/cost']) + [G[f'f{f}/cost'][-1]]                               # 50 points, padded to 51 like tester.py:204-205
        np.testing.assert_allclose(rows[b, :51], want, rtol=2e-5, err_msg=f'f{f}')
        assert rows[b, 51] == 100
        np.testing.assert_allclose(rows[b, 52], float(G[f'f{f}/ret']), rtol=1e-4)
    env = me.PBO_Env(probs[fids[0]], mo.L2L_Optimizer(argparse.Namespace(dim=10)))
    env.reset()
    h, c = torch.zeros(1, 10, dtype=torch.float64).cuda(), torch.zeros(1, 32, dtype=torch.float64).cuda()
    x = torch.zeros(1, 12, dtype=torch.float64).cuda()
    ys, done = [], False
    while not done:
        h, c = net.step(x, h, c)
        y, r, done = env.step(h[0].cpu().numpy())
        ys.append(float(y))
        x = torch.cat((h, torch.tensor([[float(y), 1.0]], dtype=torch.float64).cuda()), dim=1)
    np.testing.assert_allclose(ys, G[f'f{fids[0]}/y'], rtol=2e-5)
    assert r == 0 and len(env.optimizer.cost) == 100


# ------------------------------------------------------------------------------------------------ NRLPSO backbone (composed step, metabox_b200/nrlpso.py)
def test_nrlpso_engine_on_the_device_vs_oracle_sweep_and_dropin():
    """The composed NRLPSO step with the real evaluation kernel: every step the oracle replays the engine's own draws from the engine's
    previous state (120 steps: the pointer wraps); then a complete sweep with the tabular policy and the drop-in optimizer."""
    import argparse
    import torch
    from metabox_b200 import agent as ma, environment as me, nrlpso as mn, optimizer as mo, problem as mp, sweep
    from oracle import bbob as ob, nrlpso as on
    dim, maxFEs, fids = 10, 20000, [1, 3, 15, 21]
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    pack = mp.InstancePack([probs[f] for f in fids])
    eng = mn.NRLPSO_Engine(pack, list(range(len(fids))), maxFEs, 50, seed=12)
    cfg = on.Config(dim=dim, maxFEs=maxFEs)
    eng.init_population()
    g = np.random.RandomState(1)
    for t in range(1, 121):
        action = g.randint(0, 4, size=len(fids))
        prev = [eng.read_state(b) for b in range(len(fids))]
        j = eng.pointer
        feat, reward, done = eng.update(torch.as_tensor(action).cuda(), check_action=True)
        tp = {k: v.cpu().numpy() for k, v in eng.last_tape.items()}
        for b, f in enumerate(fids):
            I = oin[f]
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            q = prev[b]
            S = on.State(pop=q['pop'], vel=q['vel'], cost=q['cost'], pbest_pos=q['pbest_pos'], pbest_cost=q['pbest_cost'], gbest_row=int(q['gbest_row']),
                         gbest_pos=q['gbest_pos'], gbest_cost=np.float32(q['gbest_cost']), stag=q['stag'], qstate=q['qstate'], r_w=float(q['r_w']),
                         w=float(q['w']), pbest_neb_index=q['pbest_neb_index'], pbest_neb=q['pbest_neb'], gbest_neb_index=q['gbest_neb_index'],
                         gbest_neb=q['gbest_neb'], pointer=j, fes=int(q['fes']), log_index=int(q['log_index']), curve=list(q['curve'][:int(q['curve_len'])]))
            N, state, r, d = on.update(cfg, S, int(action[b]), ev, I.optimum, -5., 5., on.ArrayTape({k: v[b] for k, v in tp.items()}))
            st = eng.read_state(b)
            tag = f'step {t} lane {b} (fid {f})'
            np.testing.assert_array_equal(st['vel'], N.vel, err_msg=tag)
            np.testing.assert_array_equal(st['pop'][j], N.pop[j], err_msg=tag)
            np.testing.assert_allclose(st['cost'][j], N.cost[j], rtol=1e-6, atol=1e-10, err_msg=tag)
            if int(st['fes']) == N.fes and int(st['gbest_row']) == N.gbest_row:          # same branch decisions (a last-bit cost difference may flip a strict comparison)
                np.testing.assert_array_equal(st['pop'], N.pop, err_msg=tag)
                np.testing.assert_array_equal(st['pbest_pos'], N.pbest_pos, err_msg=tag)
                np.testing.assert_array_equal(st['stag'], N.stag, err_msg=tag)
            assert int(feat[b]) == state and int(st['log_index']) == N.log_index, tag
    qt = ma.NRLPSO_QTable(np.array([[1., 0, 0, 2], [0, 1, 0, 0], [0, 0, 1, 0], [2, 0, 0, 1]])).cuda()
    pk, pr = _pack()
    inst, _ = sweep.lane_table(len(pr), 2)
    r = sweep.gather_results(sweep.NRLPSO_Sweep(pk, inst, ma.fuse(qt), 600, 50, seed=4).run(), 1).cpu().numpy()
    assert np.all((r[:, 51] >= 600) | (r[:, 50] <= 1e-8)) and np.all(r[:, 51] <= 602) and np.all(np.diff(r[:, :51], axis=1) <= 0) and np.isfinite(r).all()
    env = me.PBO_Env(pr[0], mo.NRLPSO_Optimizer(argparse.Namespace(dim=10, maxFEs=400, n_logpoint=50, log_interval=8)))
    np.random.seed(0)
    state = env.reset()
    done, n = False, 0
    while not done:
        state, reward, done = env.step(int(np.random.randint(0, 4)))
        assert state in (0, 1, 2, 3) and reward in (2, 1, 0, -2)
        n += 1
    assert n <= 300 and env.optimizer.fes >= 400
    with pytest.raises(ValueError):
        env.step(7)


# ------------------------------------------------------------------------------------------------ LES (metabox_b200/les.py)
def test_les_engine_on_the_device_replays_reference_generations_sweep_and_dropin():
    """Golden generations of the unmodified reference (tests/golden/les.npz) through the engine with the real float64 evaluation kernel,
    then a complete sweep and the drop-in optimizer."""
    import argparse
    import torch
    from metabox_b200 import environment as me, les as ml, optimizer as mo, problem as mp, sweep
    G = np.load(os.path.join(GOLDEN, 'les.npz'))
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    pack = mp.InstancePack([probs[1], probs[15]])
    eng = ml.LES_Engine(pack, [0, 1], 20000, 50, seed=3).set_params(G['attn'], G['mlp'])
    keys = ['f1_d10', 'f15_d10']
    mu_u, zn = [], []
    for key in keys:
        np.random.seed(int(G[f'{key}/seed0']))
        mu_u.append(np.random.rand(10)); zn.append(np.random.standard_normal((16, 10)))
    eng.init_population(tape={'mu_u': np.stack(mu_u), 'normal': np.stack(zn)})
    for b, key in enumerate(keys):
        np.testing.assert_allclose(eng.cost[b].cpu().numpy(), G[f'{key}/s0/cost'], rtol=1e-10)
    for t in range(1, 41):
        zn = []
        for key in keys:
            np.random.seed(int(G[f'{key}/seed'][t - 1])); zn.append(np.random.standard_normal((16, 10)))
        eng.update(tape={'normal': np.stack(zn)})
        for b, key in enumerate(keys):
            for k in ('parents', 'mu', 'sigma', 'Pc', 'Ps', 'cost'):
                np.testing.assert_allclose(getattr(eng, k)[b].cpu().numpy(), G[f'{key}/{k}'][t - 1], rtol=1e-4, atol=1e-6, err_msg=f'{key} generation {t} {k}')
            assert int(eng.fes[b]) == int(G[f'{key}/fes'][t - 1])
    pk, pr = _pack()
    inst, _ = sweep.lane_table(len(pr), 2)
    r = sweep.gather_results(sweep.LES_Sweep(pk, inst, {'attn': G['attn'], 'mlp': G['mlp']}, 1600, 50, seed=4).run(), 1).cpu().numpy()
    assert np.all(r[:, 51] == 1600) and np.all(np.diff(r[:, :51], axis=1) <= 0) and np.isfinite(r).all() and np.all(r[:, 52] >= 0)
    opt = mo.LES_Optimizer(argparse.Namespace(dim=10, maxFEs=1600, n_logpoint=50, log_interval=32))
    env = me.PBO_Env(pr[0], opt)
    assert env.reset() is None
    gb, ret, is_end, info = env.step({'attn': G['attn'], 'mlp': G['mlp']})
    assert is_end and opt.FEs == 1600 and 0 <= ret < 1 and len(opt.cost) == 51 and abs(gb - opt.cost[-1]) <= 1e-5 * abs(gb) and info == {}


def test_gleet_trainer_with_fused_collection():
    """Trajectories sampled by the single-kernel actor: the stored log-probabilities equal the torch module's for the same (state, action),
    the packed weights follow the optimiser steps."""
    import argparse
    import torch
    from metabox_b200 import agent as ma, problem as mp, trainer as mt
    from metabox_b200.optimizer import GLEET_Engine
    cfg = argparse.Namespace(dim=10, maxFEs=2500, n_logpoint=50, log_interval=50)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    W, _ = _weights('gleet0')
    T = mt.GLEET_Trainer(cfg, tr.data, actor=ma.load_reference_weights(ma.GLEET_Actor(), W), seed=5, fused_collection=True)
    eng = GLEET_Engine(T.pack, np.arange(3, dtype=np.int32), 2500, 50, 50, seed=1)
    eng.init_population()
    seg = T.collect_segment(eng)
    with torch.no_grad():
        want = mt._GleetActorHead(T.actor).log_prob(seg['states'], seg['actions'])
    np.testing.assert_allclose(seg['logp'].cpu().numpy(), want.cpu().numpy(), rtol=2e-4, atol=2e-3)
    packed0 = T._fused.packed.clone()
    info = T.train_episodes([0, 1])
    assert info['learn_steps'] == 9 and not torch.equal(T._fused.packed, packed0)
    assert torch.equal(T._fused.packed, T.actor.pack())
