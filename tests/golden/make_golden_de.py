"""Golden sections that need the reference's agents / DE-family optimizers (imported by make_golden.py):
agents  -> agents.npz : policy weights of the shipped `agent_model/test/bbob_easy/*.pkl` agents + forward KATs
lde     -> lde.npz    : teacher-forced LDE_Optimizer trajectories (storage contract, see _refenv.F32Problem)
deddqn  -> deddqn.npz : teacher-forced DE_DDQN_Optimizer trajectories
"""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402
import torch  # noqa: E402

META = dict(numpy=np.__version__, torch=torch.__version__)
MODEL_DIR = os.path.join(_refenv.REF, 'src', 'agent_model', 'test', 'bbob_easy')


def _load_agent(name):
    import agent  # noqa: F401  (registers classes for unpickling)
    with open(os.path.join(MODEL_DIR, name + '.pkl'), 'rb') as f:
        return pickle.load(f)


def gen_agents():
    out = dict(meta=str(META))
    torch.manual_seed(0)
    # ---- RLEPSO actor (rlepso_agent.py:10-47)
    ag = _load_agent('RLEPSO_Agent')
    actor = ag._RLEPSO_Agent__actor
    for k, v in actor.state_dict().items():
        out['rlepso/' + k] = v.detach().cpu().numpy()
    xs = torch.tensor([[0.005], [0.1], [0.37], [0.8], [1.0]])
    mu = (torch.tanh(actor._Actor__mu_net(xs)) + 1.) / 2.
    sg = (torch.tanh(actor._Actor__sigma_net(xs)) + 1.) / 2. * (0.7 - 0.01) + 0.01
    out['rlepso_kat/x'] = xs.numpy(); out['rlepso_kat/mu'] = mu.detach().numpy(); out['rlepso_kat/sigma'] = sg.detach().numpy()
    # ---- DE-DDQN Q-network (de_ddqn_agent.py:21-25)
    ag = _load_agent('DE_DDQN_Agent')
    net = ag._DE_DDQN_Agent__pred_func
    for k, v in net.state_dict().items():
        out['deddqn/' + k] = v.detach().cpu().numpy()
    g = np.random.RandomState(0)
    st = torch.tensor(g.rand(16, 99), dtype=torch.float32)
    out['deddqn_kat/x'] = st.numpy(); out['deddqn_kat/q'] = net(st).detach().numpy()
    # ---- LDE policy (lde_agent.py:8-29)
    ag = _load_agent('LDE_Agent')
    net = ag._LDE_Agent__net
    for k, v in net.state_dict().items():
        out['lde/' + k] = v.detach().cpu().numpy()
    x = torch.tensor(g.rand(1, 3, 60), dtype=torch.float32)
    h = torch.tensor(g.randn(1, 3, 50) * 0.1, dtype=torch.float32)
    c = torch.tensor(g.randn(1, 3, 50) * 0.1, dtype=torch.float32)
    mu, sigma, h_, c_ = net.forward(x, h, c)
    for k, v in dict(x=x, h=h, c=c, mu=mu, sigma=sigma, h_out=h_, c_out=c_).items():
        out['lde_kat/' + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, 'agents.npz'), **out)
    print('agents done', {k: v.shape for k, v in out.items() if hasattr(v, 'shape') and k.startswith(('rlepso/', 'lde/'))})


# ------------------------------------------------------------------------------------------ LDE
LDE_CASES = [  # (key, fid or protein id, dim, maxFEs or None, n_steps)
    ('f1_d10', 1, 10, None, 5), ('f15_d10', 15, 10, None, 5), ('f8_d30', 8, 30, None, 4), ('f20_d10', 20, 10, None, 3),
    ('f102_d10', 102, 10, None, 4), ('f130_d10', 130, 10, None, 3), ('f1_d10_short', 1, 10, 400, 7),
    ('prot_1R0R_1', '1R0R_1', 12, 1000, 4),
]


def _f32r(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def _lde_get(opt):
    return dict(pop=opt._LDE_Optimizer__pop[0].astype(np.float32), fit=opt._LDE_Optimizer__fit[0].astype(np.float32),
                past_histo=opt._LDE_Optimizer__past_histo[0].copy(), gbest=np.float64(opt.gbest_cost),
                fes=np.int64(opt.fes), log_index=np.int64(opt.log_index), curve=np.array(opt.cost, dtype=np.float64))


def gen_lde():
    """Teacher-forced `LDE_Optimizer.init_population/update` (lde_optimizer.py:132-198) on F32Problem proxies."""
    import make_golden as mg
    from optimizer import LDE_Optimizer
    out = dict(meta=str(META), cases=np.array([c[0] for c in LDE_CASES]))
    ga = np.random.RandomState(777)
    for key, fid, dim, maxFEs, n_steps in LDE_CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, maxFEs, 5)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim, maxFEs)
        prob = _refenv.F32Problem(inner)
        opt = LDE_Optimizer(cfg)
        prob.reset()
        seed0 = 2000 + len(key) * 11
        np.random.seed(seed0); torch.manual_seed(seed0)
        feat = opt.init_population(prob)
        opt._LDE_Optimizer__pop = _f32r(opt._LDE_Optimizer__pop)
        for k, v in dict(fid=np.array(str(fid)), dim=np.int64(dim), maxFEs=np.int64(cfg.maxFEs),
                         n_logpoint=np.int64(cfg.n_logpoint), seed0=np.int64(seed0), n_steps=np.int64(n_steps)).items():
            out[f'{key}/{k}'] = v
        for k, v in _lde_get(opt).items():
            out[f'{key}/s0/{k}'] = v
        out[f'{key}/s0/feature'] = np.asarray(feat, dtype=np.float64)[0]
        for t in range(1, n_steps + 1):
            action = ga.rand(1, 100).astype(np.float32)
            if t == 2:
                action[0, :50] *= 1.8                     # sf > 1 never comes from the policy (clip) but drives trials out of bounds
            seed = seed0 + t
            np.random.seed(seed); torch.manual_seed(seed)
            feat, reward, done = opt.update(action, prob)
            opt._LDE_Optimizer__pop = _f32r(opt._LDE_Optimizer__pop)
            out[f'{key}/s{t}/action'] = action[0]; out[f'{key}/s{t}/seed'] = np.int64(seed)
            out[f'{key}/s{t}/feature'] = np.asarray(feat, dtype=np.float64)[0]
            out[f'{key}/s{t}/reward'] = np.float64(np.asarray(reward).reshape(-1)[0]); out[f'{key}/s{t}/done'] = np.bool_(done)
            for k, v in _lde_get(opt).items():
                out[f'{key}/s{t}/{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'lde.npz'), **out)
    print('lde done')


# ------------------------------------------------------------------------------------------ DE-DDQN
DEDDQN_CASES = [  # (key, fid or protein id, dim, maxFEs or None, n_steps)
    ('f1_d10', 1, 10, None, 330), ('f15_d10', 15, 10, None, 260), ('f9_d30', 9, 30, None, 120),
    ('f103_d10', 103, 10, None, 130), ('f1_d10_short', 1, 10, 150, 50), ('prot_1R0R_1', '1R0R_1', 12, 1000, 110),
]


def _dd_round(opt):
    for nm in ('_DE_DDQN_Optimizer__X', '_DE_DDQN_Optimizer__X_gbest', '_DE_DDQN_Optimizer__X_prebest'):
        a = getattr(opt, nm)
        np.copyto(a, _f32r(a))                                    # in place: keeps the reference's view aliasing intact


def _dd_scalars(opt):
    g = lambda n: getattr(opt, '_DE_DDQN_Optimizer__' + n)
    return np.array([g('pointer'), g('gen'), g('stagcount'), opt.fes, opt.log_index], dtype=np.int64), \
        np.array([g('c_gbest'), g('c_gworst'), g('c_prebest')], dtype=np.float64)


def gen_deddqn():
    """`DE_DDQN_Optimizer.init_population/update` (de_ddqn_optimizer.py:46-220) on F32Problem proxies, run
    continuously with its state rounded to float32 in place after every step; random actions 0..3."""
    import make_golden as mg
    from optimizer import DE_DDQN_Optimizer
    out = dict(meta=str(META), cases=np.array([c[0] for c in DEDDQN_CASES]))
    ga = np.random.RandomState(31337)
    for key, fid, dim, maxFEs, n_steps in DEDDQN_CASES:
        if isinstance(fid, str):
            inner = mg.protein_instance(fid); cfg = mg.ref_config(12, maxFEs, 5)
        else:
            inner = mg.bbob_instance(fid, dim); cfg = mg.ref_config(dim, maxFEs)
        prob = _refenv.F32Problem(inner)
        opt = DE_DDQN_Optimizer(cfg)
        prob.reset()
        seed0 = 3000 + len(key) * 13
        np.random.seed(seed0)
        feat = opt.init_population(prob)
        _dd_round(opt)
        g = lambda n: getattr(opt, '_DE_DDQN_Optimizer__' + n)
        for k, v in dict(fid=np.array(str(fid)), dim=np.int64(dim), maxFEs=np.int64(cfg.maxFEs),
                         n_logpoint=np.int64(cfg.n_logpoint), seed0=np.int64(seed0)).items():
            out[f'{key}/{k}'] = v
        out[f'{key}/X0'] = g('X').astype(np.float32); out[f'{key}/cost0'] = g('cost').astype(np.float32)
        feats, rewards, dones, ints, flts, rows, rcost, rs, acts, gb, pbx = [np.asarray(feat)], [], [], [], [], [], [], [g('r').copy()], [], [], []
        i0, f0 = _dd_scalars(opt); ints.append(i0); flts.append(f0)
        snaps = {}
        for t in range(1, n_steps + 1):
            action = int(ga.randint(0, 4))
            p = g('pointer')
            np.random.seed(seed0 + t)
            feat, reward, done = opt.update(action, prob)
            _dd_round(opt)
            i1, f1 = _dd_scalars(opt)
            acts.append(action); feats.append(np.asarray(feat)); rewards.append(reward); dones.append(done)
            ints.append(i1); flts.append(f1); rows.append(g('X')[p].astype(np.float32)); rcost.append(g('cost')[p])
            rs.append(g('r').copy()); gb.append(g('X_gbest').astype(np.float32)); pbx.append(g('X_prebest').astype(np.float32))
            if t % 50 == 0 or t == n_steps:
                snaps[t] = (g('X').astype(np.float32), g('cost').astype(np.float32))
            if done:
                break
        out[f'{key}/n_steps'] = np.int64(len(acts))
        out[f'{key}/action'] = np.array(acts, dtype=np.int64); out[f'{key}/feature'] = np.stack(feats)
        out[f'{key}/reward'] = np.array(rewards, dtype=np.float64); out[f'{key}/done'] = np.array(dones)
        out[f'{key}/ints'] = np.stack(ints); out[f'{key}/flts'] = np.stack(flts)
        out[f'{key}/row'] = np.stack(rows); out[f'{key}/row_cost'] = np.array(rcost, dtype=np.float64)
        out[f'{key}/r'] = np.stack(rs); out[f'{key}/X_gbest'] = np.stack(gb); out[f'{key}/X_prebest'] = np.stack(pbx)
        out[f'{key}/curve'] = np.array(opt.cost, dtype=np.float64)
        out[f'{key}/snap_steps'] = np.array(sorted(snaps), dtype=np.int64)
        for t, (X, c) in snaps.items():
            out[f'{key}/snapX_{t}'] = X; out[f'{key}/snapC_{t}'] = c
    np.savez_compressed(os.path.join(HERE, 'deddqn.npz'), **out)
    print('deddqn done')


SECTIONS = dict(agents=gen_agents, lde=gen_lde, deddqn=gen_deddqn)
