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


SECTIONS = dict(agents=gen_agents)
