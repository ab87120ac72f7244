"""agents_protein.npz: the RLEPSO actor of the reference's shipped protein-docking checkpoint
(`agent_model/rollout/protein_easy/RLEPSO_Agent/checkpoint20.pkl`, the policy of BASELINE config C5) + forward known answers.
Run in the build container:  python tests/golden/make_golden_protein_agent.py"""
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refenv  # noqa: E402

_refenv.activate()
import numpy as np  # noqa: E402
import torch  # noqa: E402

import agent  # noqa: E402,F401  (registers classes for unpickling)

with open(os.path.join(_refenv.REF, 'src', 'agent_model', 'rollout', 'protein_easy', 'RLEPSO_Agent', 'checkpoint20.pkl'), 'rb') as f:
    ag = pickle.load(f)
actor = ag._RLEPSO_Agent__actor
out = dict(meta=str(dict(numpy=np.__version__, torch=torch.__version__)))
for k, v in actor.state_dict().items():
    out['rlepso/' + k] = v.detach().cpu().numpy()
xs = torch.tensor([[0.1], [0.3], [0.55], [0.9], [1.0]])
mu = (torch.tanh(actor._Actor__mu_net(xs)) + 1.) / 2.
sg = (torch.tanh(actor._Actor__sigma_net(xs)) + 1.) / 2. * (0.7 - 0.01) + 0.01
out['rlepso_kat/x'] = xs.numpy(); out['rlepso_kat/mu'] = mu.detach().numpy(); out['rlepso_kat/sigma'] = sg.detach().numpy()
np.savez_compressed(os.path.join(HERE, 'agents_protein.npz'), **out)
print('agents_protein done', sum(v.size for k, v in out.items() if k.startswith('rlepso/')), 'weights')
