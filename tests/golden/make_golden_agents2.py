#!/usr/bin/env python
"""Golden weights + known-answer forwards of the reference's GLEET actor / critic and RL_PSO policy network (build container only):
python tests/golden/make_golden_agents2.py -> tests/golden/agents2.npz.

gleet/*      state_dict of `GLEET_Agent.actor` (src/agent/gleet_agent.py:307-440), checkpoint 20 of the shipped
             `agent_model/rollout/bbob_easy/GLEET_Agent`; gleet_critic/* the critic (:442-469)
gleet_kat/*  the UNMODIFIED reference actor evaluated on seeded [B, NP, 27] states: decoder logits (`only_critic=True`), mu, sigma,
             log-probability of a fixed action, critic value
gleet0/*, gleet0_kat/*   the same for checkpoint 0 (the freshly initialised actor, whose mu / sigma heads are not saturated -- the trained
             checkpoint answers mu ~ 1, sigma ~ 0.7 almost everywhere, which pins the attention stack but hardly the heads); case 'r' of
             both KAT groups uses states the reference GLEET_Optimizer itself produced (tests/golden/gleet.npz)
dedqn/*, dedqn_kat/*   Q-network of the shipped `DEDQN_Agent` (src/agent/dedqn_agent.py:13-25) + Q-values on states of reference runs
qlpso_q, rl_hpsde_q    the [4, 4] Q-tables of the shipped tabular agents
rl_pso/*     state_dict of `RL_PSO_Agent`'s PolicyNetwork (src/agent/rl_pso_agent.py:9-45) from `agent_model/test/bbob_easy`
rl_pso_kat/* mu / sigma for seeded [B, 2*dim] states (dim 10: the shipped weights are dimension-specific), plus the
             reference's out-of-range action transformation evaluated on hand-built samples
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

SRC = os.path.join(_refenv.REF, 'src', 'agent_model')


def _load(path):
    import agent  # noqa: F401  (registers the classes for unpickling)
    with open(os.path.join(SRC, path), 'rb') as f:
        return pickle.load(f)


def _gleet(out, g, grp, ckpt, real):
    ag = _load(f'rollout/bbob_easy/GLEET_Agent/checkpoint{ckpt}.pkl')
    actor, critic = ag.actor.cpu().float(), ag.critic.cpu().float()
    for k, v in actor.state_dict().items():
        out[grp + '/' + k] = v.numpy()
    for k, v in critic.state_dict().items():
        out[grp + '_critic/' + k] = v.numpy()
    for name, B, NP in (('a', 3, 100), ('b', 2, 7), ('r', len(real), real.shape[1])):
        st = g.rand(B, NP, 27).astype(np.float32)
        st[..., 0] = g.rand(B, NP) * 3.0                       # cost features are not confined to [0, 1]
        x = torch.as_tensor(real if name == 'r' else st)
        logits = actor(x, only_critic=True)
        mu = (torch.tanh(actor.mu_net(logits)) + 1.) / 2.
        sigma = (torch.tanh(actor.sigma_net(logits)) + 1.) / 2. * (actor.max_sigma - actor.min_sigma) + actor.min_sigma
        fixed = torch.as_tensor(g.rand(B, NP, 1).astype(np.float32))
        _, logp, to_critic = actor(x, fixed_action=fixed, to_critic=True)
        value = critic(to_critic)[0]
        for k, v in dict(x=x, logits=logits, mu=mu, sigma=sigma, fixed=fixed, logp=logp, value=value).items():
            out[f'{grp}_kat/{name}/{k}'] = v.numpy()
    out[grp + '_kat/max_sigma'] = np.float64(actor.max_sigma); out[grp + '_kat/min_sigma'] = np.float64(actor.min_sigma)
    print(grp, 'mu range on reference-produced states', float(out[f'{grp}_kat/r/mu'].min()), float(out[f'{grp}_kat/r/mu'].max()),
          'sigma', float(out[f'{grp}_kat/r/sigma'].min()), float(out[f'{grp}_kat/r/sigma'].max()))


def main():
    out = dict(meta=str(dict(numpy=np.__version__, torch=torch.__version__)))
    g = np.random.RandomState(5)
    with torch.no_grad():
        # ------------------------------------------------------------------ GLEET
        GS = np.load(os.path.join(HERE, 'gleet.npz'))
        real = np.stack([GS[f'{c}/s{k}/state'] for c, k in (('f1_d10', 0), ('f3_d10', 3), ('f15_d30', 1), ('f21_d10', 2))]).astype(np.float32)
        _gleet(out, g, 'gleet', 20, real)
        _gleet(out, g, 'gleet0', 0, real)
        # ------------------------------------------------------------------ RL_PSO
        ag = _load('test/bbob_easy/RL_PSO_Agent.pkl')
        net = ag._RL_PSO_Agent__nets.cpu()
        for k, v in net.state_dict().items():
            out['rl_pso/' + k] = v.numpy()
        mu_net, sg_net = net._PolicyNetwork__mu_net, net._PolicyNetwork__sigma_net
        in_dim = mu_net.state_dict()['net.layer0-linear.weight'].shape[1]
        x = torch.as_tensor(g.rand(9, in_dim).astype(np.float32) * 2 - 0.5)
        mu = (torch.tanh(mu_net(x)) + 1.) / 2.
        sigma = torch.clamp((torch.tanh(sg_net(x)) + 1.) / 2., min=net._PolicyNetwork__min_sigma, max=net._PolicyNetwork__max_sigma)
        # the reference's transformation of samples outside (0, 1) (rl_pso_agent.py:33-34), on hand-built samples
        eps = torch.as_tensor(np.array([[-3.], [-1.], [-.2], [0.], [.1], [.5], [1.], [2.5], [4.]], dtype=np.float32))
        sample = mu + sigma * eps
        filt = torch.abs(sample - 0.5) >= 0.5
        action = torch.where(filt, (sample + 3 * sigma - mu) * (1. / 6 * sigma), sample)
        for k, v in dict(x=x, mu=mu, sigma=sigma, eps=eps, action=action).items():
            out['rl_pso_kat/' + k] = v.numpy()
        out['rl_pso_kat/max_sigma'] = np.float64(net._PolicyNetwork__max_sigma); out['rl_pso_kat/min_sigma'] = np.float64(net._PolicyNetwork__min_sigma)
        # ------------------------------------------------------------------ DEDQN (dedqn_agent.py:13-25: MLP 4 -> 10 -> 10 -> 3)
        ag = _load('test/bbob_easy/DEDQN_Agent.pkl')
        dqn = ag._DEDQN_Agent__dqn.cpu()
        for k, v in dqn.state_dict().items():
            out['dedqn/' + k] = v.numpy()
        GD = np.load(os.path.join(HERE, 'dedqn.npz'))
        x = torch.as_tensor(np.concatenate([GD[f'{c}/state'][::13] for c in GD['cases']]).astype(np.float32))     # states of reference runs
        out['dedqn_kat/x'] = x.numpy(); out['dedqn_kat/q'] = dqn(x).numpy()
        # ------------------------------------------------------------------ tabular agents: the shipped Q-tables (qlpso_agent.py:17, rl_hpsde_agent.py:13)
        out['qlpso_q'] = np.asarray(_load('test/bbob_easy/QLPSO_Agent.pkl')._QLPSO_Agent__q_table, dtype=np.float64)
        out['rl_hpsde_q'] = np.asarray(_load('test/bbob_easy/RL_HPSDE_Agent.pkl')._RL_HPSDE_Agent__q_table, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'agents2.npz'), **out)
    print('rl_pso mu', mu.min().item(), mu.max().item(), 'sigma', sigma.min().item(), sigma.max().item(), 'filtered samples:', int(filt.sum()), 'of', filt.numel())


if __name__ == '__main__':
    main()
