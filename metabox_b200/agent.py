"""Agent layer: the rollout-time policies of the MetaBBO agents whose backbones are built here, as plain torch modules
running on the same CUDA stream as the step kernels (north star: "the small src/agent policy MLP/LSTM
left as a torch call on the same stream").  The learning updates live in trainer.py (LDE REINFORCE, RLEPSO PPO).

Reference:  MLP builder `src/agent/networks.py:4-26`; RLEPSO `Actor.forward` `rlepso_agent.py:22-47`,
rollout loop `:294-303`; DE-DDQN net `de_ddqn_agent.py:21-25`, greedy action `:59-68`, rollout `:108-117`;
LDE `PolicyNet` `lde_agent.py:8-29`, rollout `:147-163`; GLEET attention actor / critic `gleet_agent.py:307-469` on
`networks.py:47-356`, rollout `:293-303`; RL_PSO `PolicyNetwork` `rl_pso_agent.py:9-45`, rollout `:116-124`;
QLPSO Q-table `qlpso_agent.py:17-48`.

Module / parameter names mirror the reference so that a `state_dict` extracted from a reference pickle
(see tests/golden/make_golden_de.py: `agents.npz`) loads with `load_reference_weights`.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


class MLP(nn.Module):
    """networks.py:4-26 -- list of {'in','out','drop_out','activation'} dicts."""

    def __init__(self, config):
        super().__init__()
        self.net = nn.Sequential()
        self.net_config = config
        for k, lc in enumerate(config):
            self.net.add_module(f'layer{k}-linear', nn.Linear(lc['in'], lc['out']))
            self.net.add_module(f'layer{k}-drop_out', nn.Dropout(lc['drop_out']))
            if lc['activation'] != 'None':
                self.net.add_module(f'layer{k}-activation', getattr(nn, lc['activation'])())

    def forward(self, x):
        return self.net(x)


def _mlp_cfg(sizes, last_act='None'):
    cfg = []
    for k in range(len(sizes) - 1):
        act = 'ReLU' if k < len(sizes) - 2 else last_act
        cfg.append({'in': sizes[k], 'out': sizes[k + 1], 'drop_out': 0, 'activation': act})
    return cfg


def load_reference_weights(module: nn.Module, weights: dict, prefix=''):
    """weights: {reference state_dict key: array}.  Name-mangled private attributes of the reference
    (`_Actor__mu_net`) are mapped onto this module's public names (`mu_net`); a module may list further
    (reference substring -> own substring) renames in `reference_renames`."""
    own = module.state_dict()
    mapped = {}
    for k, v in weights.items():
        if prefix and not k.startswith(prefix):
            continue
        kk = k[len(prefix):]
        parts = [p.split('__', 1)[1] if p.startswith('_') and '__' in p[1:] else p for p in kk.split('.')]
        kk = '.'.join(parts)
        for a, b in getattr(module, 'reference_renames', ()):
            kk = kk.replace(a, b)
        if kk in own:
            mapped[kk] = torch.as_tensor(np.asarray(v)).to(own[kk].dtype).reshape(own[kk].shape)
    missing = set(own) - set(mapped)
    if missing:
        raise KeyError(f'reference weights missing for {sorted(missing)}')
    module.load_state_dict(mapped)
    return module


# ------------------------------------------------------------------------------------------ RLEPSO
class RLEPSO_Actor(nn.Module):
    """Two MLPs 1->64->32->35 (mu, sigma); action = clamp(Normal(mu, sigma).sample(), 0, 1)."""
    feature_dim, action_dim, max_sigma, min_sigma = 1, 35, 0.7, 0.01
    takes_noise = True          # forward(x, noise=...) accepts externally drawn standard-normal noise

    def __init__(self):
        super().__init__()
        cfg = _mlp_cfg([self.feature_dim, 64, 32, self.action_dim])
        self.mu_net = MLP(cfg)
        self.sigma_net = MLP(cfg)

    def distribution(self, x):
        mu = (torch.tanh(self.mu_net(x)) + 1.) / 2.
        sigma = (torch.tanh(self.sigma_net(x)) + 1.) / 2. * (self.max_sigma - self.min_sigma) + self.min_sigma
        return mu, sigma

    def log_prob(self, x, action):
        """sum over the 35 coefficients of log N(action; mu, sigma) (rlepso_agent.py:26-35, `fixed_action` path); differentiable."""
        mu, sigma = self.distribution(x)
        return torch.distributions.Normal(mu, sigma).log_prob(action).sum(-1)

    @torch.no_grad()
    def forward(self, x, noise=None):
        """x [..., 1] -> action [..., 35].  `noise` (standard normal, same shape as the action) can be
        injected for deterministic tests; otherwise torch.normal on the current device/stream."""
        mu, sigma = self.distribution(x)
        eps = torch.randn_like(mu) if noise is None else noise
        return torch.clamp(mu + sigma * eps, min=0, max=1)


    @torch.no_grad()
    def pack(self, out=None):
        """All parameters as ONE float32 vector in the layout `mbx_rlepso_policy` reads (csrc/mbx_rlepso_policy.cuh): per network (mu, sigma)
        L0 w [64] | b [64] | L1 W^T [64][32] | b [32] | L2 W^T [32][35] | b [35]."""
        parts = []
        for net in (self.mu_net, self.sigma_net):
            l0, l1, l2 = [m for m in net.net if isinstance(m, nn.Linear)]
            parts += [l0.weight[:, 0], l0.bias, l1.weight.t(), l1.bias, l2.weight.t(), l2.bias]
        flat = torch.cat([t.detach().to(torch.float32).reshape(-1) for t in parts])
        if out is None:
            return flat.contiguous()
        out.copy_(flat)
        return out


class RLEPSO_FusedActor:
    """`RLEPSO_Actor`'s rollout forward as ONE CUDA kernel per generation (C ABI `mbx_rlepso_policy`, csrc/mbx_rlepso_policy.cu) instead of
    ~20 torch launches; same call convention (`policy(state [B,1], noise=None) -> action [B,35]`, `takes_noise`, `action_dim`), so it
    plugs into `sweep.RLEPSO_Sweep`.  `refresh()` re-packs the weights in place after an optimiser step on `actor`."""
    takes_noise = True
    fused_action = True
    action_dim = 35

    def __init__(self, actor: 'RLEPSO_Actor'):
        from . import _lib
        self.actor = actor
        self.packed = actor.pack()
        n = _lib.lib().mbx_rlepso_policy_weight_count()
        if self.packed.numel() != n:
            raise _lib.MbxError(f'RLEPSO_Actor.pack() produced {self.packed.numel()} weights, libmbx.so expects {n}')
        if not self.packed.is_cuda:
            raise _lib.MbxError('RLEPSO_FusedActor needs the actor on a CUDA device (there is no CPU path)')

    def refresh(self):
        self.actor.pack(out=self.packed)
        return self

    @torch.no_grad()
    def __call__(self, x, noise=None, outputs=False):
        """-> action [B,35]; with `outputs=True` -> (action, mu [B,35], sigma [B,35])."""
        from . import _lib
        if x.dtype != torch.float32 or x.device != self.packed.device or x.numel() != x.shape[0]:
            raise ValueError(f'state must be float32 [B] or [B, 1] on {self.packed.device}: got {x.dtype} {tuple(x.shape)} on {x.device}')
        x = x.contiguous()
        B = x.shape[0]
        if noise is not None:
            noise = noise.to(torch.float32).reshape(B, 35).contiguous()
        action = torch.empty(B, 35, dtype=torch.float32, device=x.device)
        mu, sigma = (torch.empty_like(action), torch.empty_like(action)) if outputs else (None, None)
        with torch.cuda.device(x.device):
            rc = _lib.lib().mbx_rlepso_policy(_lib.ptr(x), _lib.ptr(self.packed), self.packed.numel(), _lib.ptr(noise), _lib.ptr(action),
                                              _lib.ptr(mu), _lib.ptr(sigma), B, _lib.stream_ptr())
        _lib.check(rc, 'mbx_rlepso_policy')
        return (action, mu, sigma) if outputs else action


class RLEPSO_Critic(nn.Module):
    """Value head 1 -> 16 -> 8 -> 1 (rlepso_agent.py:49-61); forward(x [..., 1]) -> value [...]."""

    def __init__(self):
        super().__init__()
        self.value_head = MLP(_mlp_cfg([RLEPSO_Actor.feature_dim, 16, 8, 1]))

    def forward(self, x):
        return self.value_head(x).squeeze(-1)


class RLEPSO_Agent:
    """Rollout side of `RLEPSO_Agent` (rlepso_agent.py:294-303) for one env or a batched engine."""

    def __init__(self, config=None, device='cuda'):
        self.device = torch.device(device)
        self.actor = RLEPSO_Actor().to(self.device).eval()

    def get_action(self, state):
        state = torch.as_tensor(np.asarray(state), dtype=torch.float32).to(self.device)
        return self.actor(state).cpu().numpy()

    def rollout_episode(self, env):
        is_done = False
        state = env.reset()
        R = 0
        while not is_done:
            action = self.get_action(state)
            state, reward, is_done = env.step(action)
            R += reward
        return {'cost': env.optimizer.cost, 'fes': env.optimizer.fes, 'return': R}


# ------------------------------------------------------------------------------------------ DE-DDQN
class DE_DDQN_Net(MLP):
    """99 -> 100 -> 100 -> 100 -> 100 -> 4 (de_ddqn_agent.py:21-25); greedy action = argmax Q."""

    def __init__(self):
        super().__init__(_mlp_cfg([99, 100, 100, 100, 100, 4]))

    @torch.no_grad()
    def act(self, state):
        return torch.argmax(self.forward(state), dim=-1)

    @torch.no_grad()
    def pack(self, out=None):
        """All parameters as ONE float32 vector in the layout `mbx_deddqn_policy` reads (csrc/mbx_deddqn_policy.cuh): the four hidden
        layers as W^T [in][100] | bias, the output layer as W [4][100] | bias."""
        lin = [m for m in self.net if isinstance(m, nn.Linear)]
        parts = [t for l in lin[:-1] for t in (l.weight.t(), l.bias)] + [lin[-1].weight, lin[-1].bias]
        flat = torch.cat([t.detach().to(torch.float32).reshape(-1) for t in parts])
        if out is None:
            return flat.contiguous()
        out.copy_(flat)
        return out


class DE_DDQN_FusedNet:
    """`DE_DDQN_Net`'s greedy action as ONE CUDA kernel per environment step (C ABI `mbx_deddqn_policy`, csrc/mbx_deddqn_policy.cu) instead
    of ~12 torch launches: `policy(state [B,99]) -> action [B] int32`; `sweep.DE_DDQN_Sweep` uses it when handed one.  `refresh()` re-packs
    the weights in place after an optimiser step on `net` (graphs that captured the launch see the new values)."""
    fused_action = True

    def __init__(self, net: 'DE_DDQN_Net'):
        from . import _lib
        self.net = net
        self.packed = net.pack()
        n = _lib.lib().mbx_deddqn_policy_weight_count()
        if self.packed.numel() != n:
            raise _lib.MbxError(f'DE_DDQN_Net.pack() produced {self.packed.numel()} weights, libmbx.so expects {n}')
        if not self.packed.is_cuda:
            raise _lib.MbxError('DE_DDQN_FusedNet needs the network on a CUDA device (there is no CPU path)')

    def refresh(self):
        self.net.pack(out=self.packed)
        return self

    @torch.no_grad()
    def __call__(self, state, q=False):
        """-> action [B] int32; with `q=True` -> (action, Q [B,4])."""
        from . import _lib
        if state.dim() != 2 or state.shape[1] != 99 or state.dtype != torch.float32 or state.device != self.packed.device:
            raise ValueError(f'state must be float32 [B, 99] on {self.packed.device}: got {state.dtype} {tuple(state.shape)} on {state.device}')
        state = state.contiguous()
        B = state.shape[0]
        action = torch.empty(B, dtype=torch.int32, device=state.device)
        qv = torch.empty(B, 4, dtype=torch.float32, device=state.device) if q else None
        with torch.cuda.device(state.device):
            rc = _lib.lib().mbx_deddqn_policy(_lib.ptr(state), _lib.ptr(self.packed), self.packed.numel(), _lib.ptr(action), _lib.ptr(qv), B,
                                              _lib.stream_ptr())
        _lib.check(rc, 'mbx_deddqn_policy')
        return (action, qv) if q else action


class DEDQN_Net(MLP):
    """Q-network of DEDQN_Agent: 4 -> 10 -> 10 -> 3 (dedqn_agent.py:13-15); greedy rollout action = argmax Q (:44-53, :87-96)."""

    def __init__(self):
        super().__init__(_mlp_cfg([4, 10, 10, 3]))

    @torch.no_grad()
    def act(self, state):
        return torch.argmax(self.forward(state), dim=-1)


# ------------------------------------------------------------------------------------------ LDE
class LDE_PolicyNet(nn.Module):
    """LSTM(NP+10 -> 50) + Linear mu/sigma heads -> 2*NP (lde_agent.py:8-29)."""
    takes_noise = True          # step(x, h, c, noise=...) accepts externally drawn standard-normal noise

    def __init__(self, NP=50, cell=50, bins=5):
        super().__init__()
        self.NP, self.cell = NP, cell
        self.lstm = nn.LSTM(input_size=NP + 2 * bins, hidden_size=cell, num_layers=1)
        self.mu = nn.Linear(cell, 2 * NP)
        self.sigma = nn.Linear(cell, 2 * NP)

    def forward(self, x, h, c):
        out, (h_, c_) = self.lstm(x, (h, c))
        return self.mu(out), torch.sigmoid(self.sigma(out)), h_, c_

    @torch.no_grad()
    def sampler(self, x, h, c, noise=None):
        """x [1,B,NP+10] -> action [B,2NP] clipped to [0,1], new (h,c)."""
        mu, sigma, h_, c_ = self.forward(x, h, c)
        eps = torch.randn_like(mu) if noise is None else noise
        return torch.clip(mu + sigma * eps, 0, 1)[0], h_, c_

    def lstm_cell(self, x, h, c):
        """One LSTM step written out (gate order i, f, g, o like torch.nn.LSTM) so that the rollout loop is plain
        matmuls on the current stream (CUDA-graph capturable, no cuDNN RNN descriptor).  x [B,NP+10], h/c [B,cell]
        -> mu [B,2NP], sigma [B,2NP], h', c'."""
        L = self.lstm
        gates = torch.addmm(L.bias_ih_l0 + L.bias_hh_l0, x, L.weight_ih_l0.t()) + h @ L.weight_hh_l0.t()
        i, f, g, o = gates.chunk(4, dim=-1)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        return self.mu(h2), torch.sigmoid(self.sigma(h2)), h2, c2

    @torch.no_grad()
    def step(self, x, h, c, noise=None):
        """Batched `sampler` for B lanes: action [B,2NP] = clip(Normal(mu, sigma).sample(), 0, 1), new (h, c)."""
        mu, sigma, h2, c2 = self.lstm_cell(x, h, c)
        eps = torch.randn_like(mu) if noise is None else noise
        return torch.clip(mu + sigma * eps, 0, 1), h2, c2


# ------------------------------------------------------------------------------------------ L2L
class L2L_PolicyNet(nn.Module):
    """nn.LSTM(dim + 2 -> 32, proj_size = dim) in float64 (l2l_agent.py:20-21); `step` is the cell written out (gate order i, f, g, o;
    h' = W_hr (o * tanh(c'))) so that the 100-step episode is plain matmuls on the current stream."""

    def __init__(self, dim, hidden=32):
        super().__init__()
        self.dim, self.hidden = int(dim), int(hidden)
        self.net = nn.LSTM(input_size=self.dim + 2, hidden_size=self.hidden, proj_size=self.dim, dtype=torch.float64)

    @torch.no_grad()
    def step(self, x, h, c):
        """x [B, dim+2], h [B, dim], c [B, 32] (float64) -> raw action = h' [B, dim], c'."""
        L = self.net
        gates = torch.addmm(L.bias_ih_l0 + L.bias_hh_l0, x, L.weight_ih_l0.t()) + h @ L.weight_hh_l0.t()
        i, f, g, o = gates.chunk(4, dim=-1)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        return (torch.sigmoid(o) * torch.tanh(c2)) @ L.weight_hr_l0.t(), c2


# ------------------------------------------------------------------------------------------ QLPSO
class QLPSO_QTable(nn.Module):
    """Tabular policy of QLPSO_Agent (qlpso_agent.py:17-37): a [4, 4] Q-table; the action is sampled from
    softmax(Q[state]) (`__get_action`, :33-36).  `forward(q_state [B] int, u=None) -> action [B] int32`; `u` are uniforms
    in [0, 1) (inverse-CDF sampling, so that a sweep can draw them from its own generator under CUDA-graph capture).
    `td_update` is the reference's learning rule (:47-48) for batches of transitions, applied in order."""
    takes_uniform = True

    def __init__(self, q_table=None, n_states=4, n_actions=4):
        super().__init__()
        q = torch.zeros(n_states, n_actions, dtype=torch.float64) if q_table is None else torch.as_tensor(q_table, dtype=torch.float64)
        self.register_buffer('q_table', q.clone())

    @torch.no_grad()
    def forward(self, state, u=None):
        prob = torch.softmax(self.q_table, dim=1)[state.long()]                       # exp(q) / exp(q).sum()  (:34-35)
        if u is None:
            u = torch.rand(state.shape[0], device=state.device, dtype=torch.float32)
        cdf = torch.cumsum(prob, dim=1)
        a = (u.to(cdf.dtype)[:, None] >= cdf).sum(dim=1).clamp(max=self.q_table.shape[1] - 1)
        return a.to(torch.int32)

    @torch.no_grad()
    def td_update(self, state, action, reward, next_state, alpha, gamma=0.8):
        """Q[s][a] += alpha * (r + gamma * max Q[s'] - Q[s][a]) for each transition in turn (qlpso_agent.py:47-48)."""
        q = self.q_table
        for s, a, r, n in zip(state.tolist(), action.tolist(), reward.tolist(), next_state.tolist()):
            q[s, a] += alpha * (r + gamma * q[n].max() - q[s, a])


NRLPSO_QTable = QLPSO_QTable            # NRLPSO_Agent (nrlpso_agent.py:13,22-25,36-37): the same tabular softmax policy (gamma 0.8, alpha 1)
RL_HPSDE_QTable = QLPSO_QTable          # RL_HPSDE_Agent (rl_hpsde_agent.py:13,27-30,42-43) is the same tabular softmax policy over 4 states x 4 actions (gamma 0.5)


# ------------------------------------------------------------------------------------------ GLEET
class _SwarmNorm(nn.Module):
    """The reference's 'layer' normalisation (networks.py:68-72): every lane is standardised over ALL of its [NP, E] entries
    (unbiased variance, eps 1e-5); no affine parameters."""

    def forward(self, x):
        m = x.mean(dim=(1, 2), keepdim=True)
        return (x - m) * torch.rsqrt(x.var(dim=(1, 2), keepdim=True) + 1e-05)


class GLEET_AttentionBlock(nn.Module):
    """One `MultiHeadEncoder` (networks.py:256-341): multi-head attention of the queries `q` over the particles `h` + residual on
    `h` + swarm norm, then a 16 -> 16 -> 16 ReLU feed-forward + residual + swarm norm.  Parameters keep the reference's
    [heads, in, key] layout (so its checkpoints load); the forward folds the heads into ONE [E, heads*key] projection per
    operand and lets `scaled_dot_product_attention` do softmax(QK^T / sqrt(key)) V without materialising [B, heads, NP, NP]
    in HBM when a fused backend takes it."""

    def __init__(self, n_heads=4, embed_dim=16, ff_hidden=16):
        super().__init__()
        kd = embed_dim // n_heads
        self.n_heads, self.embed_dim, self.key_dim = n_heads, embed_dim, kd
        self.W_query = nn.Parameter(torch.rand(n_heads, embed_dim, kd))
        self.W_key = nn.Parameter(torch.rand(n_heads, embed_dim, kd))
        self.W_val = nn.Parameter(torch.rand(n_heads, embed_dim, kd))
        self.W_out = nn.Parameter(torch.rand(n_heads, kd, embed_dim))
        self.ff_in = nn.Linear(embed_dim, ff_hidden)
        self.ff_out = nn.Linear(ff_hidden, embed_dim)
        self.norm = _SwarmNorm()

    def _project(self, x, W):                                   # [B, N, E] x [H, E, K] -> [B, H, N, K]
        B, N, E = x.shape
        return (x @ W.permute(1, 0, 2).reshape(E, -1)).view(B, N, self.n_heads, -1).transpose(1, 2)

    def forward(self, h, q=None):
        q = h if q is None else q
        B, N, E = q.shape
        heads = nn.functional.scaled_dot_product_attention(self._project(q, self.W_query), self._project(h, self.W_key),
                                                           self._project(h, self.W_val))        # scale = 1/sqrt(key_dim)
        att = heads.transpose(1, 2).reshape(B, N, -1) @ self.W_out.reshape(-1, E)
        x = self.norm(att + h)
        return self.norm(self.ff_out(torch.relu(self.ff_in(x))) + x)


_GLEET_RENAMES = (('.0.MHA_sublayer.MHA.', '.'), ('.0.FFandNorm_sublayer.FF.0.', '.ff_in.'), ('.0.FFandNorm_sublayer.FF.2.', '.ff_out.'),
                  ('embedder.embedder.', 'embedder.'), ('embedder_for_decoder.embedder.', 'embedder_for_decoder.'))


class GLEET_Actor(nn.Module):
    """gleet_agent.py:307-440.  state [B, NP, 27] = 9 population features | 9 exploration (pbest) | 9 exploitation (gbest) features:
    encoder block over the embedded population features; the two 9-feature groups share the SAME embedder, are concatenated and
    embedded again, and query the encoder output in the decoder block; per-particle mu / sigma heads 16 -> 32 -> 8 -> 1 (LeakyReLU).
    action [B, NP] = clamp(Normal(mu, sigma).sample(), 0, 1)."""
    node_dim, embedding_dim, max_sigma, min_sigma = 9, 16, 0.7, 0.01
    takes_noise = True
    reference_renames = _GLEET_RENAMES

    def __init__(self):
        super().__init__()
        E = self.embedding_dim
        self.embedder = nn.Linear(self.node_dim, E, bias=False)
        self.encoder = GLEET_AttentionBlock(4, E, 16)
        self.embedder_for_decoder = nn.Linear(2 * E, E, bias=False)
        self.decoder = GLEET_AttentionBlock(4, E, 16)
        cfg = [{'in': E, 'out': 32, 'drop_out': 0, 'activation': 'LeakyReLU'}, {'in': 32, 'out': 8, 'drop_out': 0, 'activation': 'LeakyReLU'},
               {'in': 8, 'out': 1, 'drop_out': 0, 'activation': 'None'}]
        self.mu_net = MLP(cfg)
        self.sigma_net = MLP(cfg)

    def logits(self, x):
        """[B, NP, 27] -> decoder output [B, NP, 16] (what the reference returns for `only_critic=True` and hands to the critic)."""
        nd = self.node_dim
        B, N, _ = x.shape
        em = self.embedder(x.reshape(B, N, 3, nd))                          # one GEMM for the three feature groups
        enc = self.encoder(em[:, :, 0])
        return self.decoder(enc, self.embedder_for_decoder(em[:, :, 1:].reshape(B, N, -1)))

    def distribution(self, x):
        lg = self.logits(x)
        mu = (torch.tanh(self.mu_net(lg)) + 1.) / 2.
        sigma = (torch.tanh(self.sigma_net(lg)) + 1.) / 2. * (self.max_sigma - self.min_sigma) + self.min_sigma
        return mu.squeeze(-1), sigma.squeeze(-1), lg

    def log_prob(self, x, action):
        """Joint log-probability of the swarm's action [B, NP] (summed over the particles, :425-427); differentiable."""
        mu, sigma, _ = self.distribution(x)
        return torch.distributions.Normal(mu, sigma).log_prob(action).sum(-1)

    @torch.no_grad()
    def forward(self, x, noise=None):
        """x [B, NP, 27] -> action [B, NP]; `noise` = externally drawn standard normals of the action's shape."""
        mu, sigma, _ = self.distribution(x)
        eps = torch.randn_like(mu) if noise is None else noise
        return torch.clamp(mu + sigma * eps, min=0, max=1)

    @torch.no_grad()
    def pack(self, out=None):
        """All parameters as ONE float32 vector in the layout `mbx_gleet_policy` reads (metabox_b200/csrc/mbx_gleet_policy.cuh):
        embedder^T | encoder block | embedder_for_decoder | decoder block | mu head | sigma head, the per-head attention matrices
        folded into [16 in][head*4 + k] / [head*4 + k][16 out], every piece starting on a multiple of 4 floats."""
        def block(b):
            fold = lambda W: W.permute(1, 0, 2).reshape(self.embedding_dim, -1)
            return [fold(b.W_query), fold(b.W_key), fold(b.W_val), b.W_out.reshape(-1, self.embedding_dim),
                    b.ff_in.weight, b.ff_in.bias, b.ff_out.weight, b.ff_out.bias]

        def head(m):
            lin = [mod for mod in m.net if isinstance(mod, nn.Linear)]
            return [t for l in lin for t in (l.weight, l.bias)] + [torch.zeros(3, device=lin[0].weight.device)]

        parts = [self.embedder.weight.t()] + block(self.encoder) + [self.embedder_for_decoder.weight] + block(self.decoder) \
            + head(self.mu_net) + head(self.sigma_net)
        flat = torch.cat([t.detach().to(torch.float32).reshape(-1) for t in parts])
        if out is None:
            return flat.contiguous()
        out.copy_(flat)
        return out


class GLEET_FusedActor:
    """`GLEET_Actor`'s rollout forward as ONE CUDA kernel per generation (C ABI `mbx_gleet_policy`, csrc/mbx_gleet_policy.cu): one CTA
    per swarm, keys / values and weights in shared memory.  Same call convention as the torch module
    (`policy(state [B,NP,27], noise=None) -> action [B,NP]`, `takes_noise`), so it plugs into `sweep.GLEET_Sweep` / `tester`.  The packed
    weights are a device buffer the kernel reads on every launch: after an optimiser step on `actor`, call `refresh()` (an in-place
    copy, so CUDA graphs that captured the launch see the new weights)."""
    takes_noise = True
    fused_action = True

    def __init__(self, actor: 'GLEET_Actor'):
        from . import _lib
        self.actor = actor
        self.packed = actor.pack()
        n = _lib.lib().mbx_gleet_policy_weight_count()
        if self.packed.numel() != n:
            raise _lib.MbxError(f'GLEET_Actor.pack() produced {self.packed.numel()} weights, libmbx.so expects {n}')
        if not self.packed.is_cuda:
            raise _lib.MbxError('GLEET_FusedActor needs the actor on a CUDA device (there is no CPU path)')

    def refresh(self):
        self.actor.pack(out=self.packed)
        return self

    @torch.no_grad()
    def __call__(self, x, noise=None, outputs=False):
        """-> action [B,NP]; with `outputs=True` -> (action, mu [B,NP], sigma [B,NP], decoder logits [B,NP,16])."""
        from . import _lib
        if x.dim() != 3 or x.shape[2] != 27 or x.dtype != torch.float32 or x.device != self.packed.device:
            raise ValueError(f'state must be float32 [B, NP, 27] on {self.packed.device}: got {x.dtype} {tuple(x.shape)} on {x.device}')
        x = x.contiguous()
        B, NP, _ = x.shape
        if noise is not None:
            noise = noise.to(torch.float32).reshape(B, NP).contiguous()
        action = torch.empty(B, NP, dtype=torch.float32, device=x.device)
        mu = sigma = logits = None
        if outputs:
            mu, sigma = torch.empty_like(action), torch.empty_like(action)
            logits = torch.empty(B, NP, 16, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().mbx_gleet_policy(_lib.ptr(x), _lib.ptr(self.packed), self.packed.numel(), _lib.ptr(noise), _lib.ptr(action),
                                             _lib.ptr(mu), _lib.ptr(sigma), _lib.ptr(logits), B, NP, _lib.stream_ptr())
        _lib.check(rc, 'mbx_gleet_policy')
        return (action, mu, sigma, logits) if outputs else action


class GLEET_Critic(nn.Module):
    """gleet_agent.py:442-469: value head 16 -> 32 -> 16 -> 1 (LeakyReLU) on the particle-mean of the decoder output."""

    def __init__(self):
        super().__init__()
        self.value_head = MLP([{'in': 16, 'out': 32, 'drop_out': 0, 'activation': 'LeakyReLU'},
                               {'in': 32, 'out': 16, 'drop_out': 0, 'activation': 'LeakyReLU'},
                               {'in': 16, 'out': 1, 'drop_out': 0, 'activation': 'None'}])

    def forward(self, logits):
        return self.value_head(logits.mean(dim=-2)).squeeze(-1)


# ------------------------------------------------------------------------------------------ RL_PSO
class RL_PSO_PolicyNet(nn.Module):
    """rl_pso_agent.py:9-45: mu / sigma MLPs 2*dim -> 32 -> 8 -> 1 (ReLU); sigma = clamp((tanh+1)/2, 0.01, 0.7); a sample `a` with
    |a - 0.5| >= 0.5 is replaced by (a + 3 sigma - mu) * (sigma / 6) (the reference's expression, operator precedence included)."""
    max_sigma, min_sigma = 0.7, 0.01
    takes_noise = True

    def __init__(self, dim):
        super().__init__()
        self.dim = int(dim)
        cfg = _mlp_cfg([2 * self.dim, 32, 8, 1])
        self.mu_net = MLP(cfg)
        self.sigma_net = MLP(cfg)

    def distribution(self, x):
        mu = (torch.tanh(self.mu_net(x)) + 1.) / 2.
        sigma = torch.clamp((torch.tanh(self.sigma_net(x)) + 1.) / 2., min=self.min_sigma, max=self.max_sigma)
        return mu, sigma

    @torch.no_grad()
    def forward(self, x, noise=None):
        """x [B, 2*dim] -> action [B, 1]."""
        mu, sigma = self.distribution(x)
        a = mu + sigma * (torch.randn_like(mu) if noise is None else noise)
        return torch.where(torch.abs(a - 0.5) >= 0.5, (a + 3 * sigma - mu) * (1. / 6 * sigma), a)

    @torch.no_grad()
    def pack(self, out=None):
        """All parameters as ONE float32 vector in the layout `mbx_rl_pso_policy` reads (csrc/mbx_pso_policies.cuh): per network (mu, sigma)
        L0 W^T [2*dim][32] | b [32] | L1 W [8][32] | b [8] | L2 w [8] | b [1]."""
        parts = []
        for net in (self.mu_net, self.sigma_net):
            l0, l1, l2 = [m for m in net.net if isinstance(m, nn.Linear)]
            parts += [l0.weight.t(), l0.bias, l1.weight, l1.bias, l2.weight, l2.bias]
        flat = torch.cat([t.detach().to(torch.float32).reshape(-1) for t in parts])
        if out is None:
            return flat.contiguous()
        out.copy_(flat)
        return out


class RL_PSO_FusedPolicy:
    """`RL_PSO_PolicyNet`'s rollout forward as ONE CUDA kernel per environment step (C ABI `mbx_rl_pso_policy`, csrc/mbx_pso_policies.cu):
    `policy(state [B, 2*dim], noise=None) -> action [B, 1]`, `takes_noise`; plugs into `sweep.RL_PSO_Sweep`.  `refresh()` re-packs."""
    takes_noise = True
    fused_action = True

    def __init__(self, net: 'RL_PSO_PolicyNet'):
        from . import _lib
        self.net, self.n_in = net, 2 * net.dim
        self.packed = net.pack()
        n = _lib.lib().mbx_rl_pso_policy_weight_count(self.n_in)
        if self.packed.numel() != n:
            raise _lib.MbxError(f'RL_PSO_PolicyNet.pack() produced {self.packed.numel()} weights, libmbx.so expects {n}')
        if not self.packed.is_cuda:
            raise _lib.MbxError('RL_PSO_FusedPolicy needs the network on a CUDA device (there is no CPU path)')

    def refresh(self):
        self.net.pack(out=self.packed)
        return self

    @torch.no_grad()
    def __call__(self, x, noise=None, outputs=False):
        from . import _lib
        if x.dim() != 2 or x.shape[1] != self.n_in or x.dtype != torch.float32 or x.device != self.packed.device:
            raise ValueError(f'state must be float32 [B, {self.n_in}] on {self.packed.device}: got {x.dtype} {tuple(x.shape)} on {x.device}')
        x = x.contiguous()
        B = x.shape[0]
        if noise is not None:
            noise = noise.to(torch.float32).reshape(B).contiguous()
        action = torch.empty(B, 1, dtype=torch.float32, device=x.device)
        mu, sigma = (torch.empty_like(action), torch.empty_like(action)) if outputs else (None, None)
        with torch.cuda.device(x.device):
            rc = _lib.lib().mbx_rl_pso_policy(_lib.ptr(x), _lib.ptr(self.packed), self.packed.numel(), _lib.ptr(noise), _lib.ptr(action),
                                              _lib.ptr(mu), _lib.ptr(sigma), B, self.n_in, _lib.stream_ptr())
        _lib.check(rc, 'mbx_rl_pso_policy')
        return (action, mu, sigma) if outputs else action


class QLPSO_FusedQTable:
    """`QLPSO_QTable`'s sampling as ONE CUDA kernel per environment step (C ABI `mbx_qlpso_policy`): `policy(q_state [B] int32, u=None) ->
    action [B] int32`, `takes_uniform`; plugs into `sweep.QLPSO_Sweep`.  The kernel reads the module's live float64 Q-table buffer, so
    `td_update` needs no refresh."""
    takes_uniform = True
    fused_action = True

    def __init__(self, qtable: 'QLPSO_QTable'):
        from . import _lib
        _lib.lib()
        if not qtable.q_table.is_cuda or tuple(qtable.q_table.shape) != (4, 4):
            raise _lib.MbxError('QLPSO_FusedQTable needs a [4, 4] Q-table on a CUDA device (there is no CPU path)')
        self.qtable = qtable

    @torch.no_grad()
    def __call__(self, state, u=None):
        from . import _lib
        q = self.qtable.q_table
        if state.dim() != 1 or state.device != q.device:
            raise ValueError(f'q_state must be [B] on {q.device}: got {tuple(state.shape)} on {state.device}')
        state = state.to(torch.int32).contiguous()
        B = state.shape[0]
        u = torch.rand(B, device=q.device, dtype=torch.float32) if u is None else u.to(torch.float32).reshape(B).contiguous()
        action = torch.empty(B, dtype=torch.int32, device=q.device)
        with torch.cuda.device(q.device):
            rc = _lib.lib().mbx_qlpso_policy(_lib.ptr(state), _lib.ptr(q), _lib.ptr(u), _lib.ptr(action), B, _lib.stream_ptr())
        _lib.check(rc, 'mbx_qlpso_policy')
        return action


def fuse(policy):
    """The single-kernel twin of a rollout policy when the library has one (RLEPSO actor, DE-DDQN Q-network, GLEET actor, RL_PSO policy
    network, QLPSO Q-table), else the policy itself.  The module must already live on its CUDA device."""
    table = ((RLEPSO_Actor, RLEPSO_FusedActor), (DE_DDQN_Net, DE_DDQN_FusedNet), (GLEET_Actor, GLEET_FusedActor),
             (RL_PSO_PolicyNet, RL_PSO_FusedPolicy), (QLPSO_QTable, QLPSO_FusedQTable))
    for cls, wrapper in table:
        if isinstance(policy, cls):
            return wrapper(policy)
    return policy
