import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from metabox_b200 import agent as ma, optimizer as mo, problem as mp
dim = 30
tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
probs = tr.data + te.data
pack = mp.InstancePack(probs)
lanes = np.repeat(np.arange(len(probs)), 4)
eng = mo.GLEET_Engine(pack, lanes, 60000, 50, seed=1)
st = eng.init_population()
torch.manual_seed(0)
actor = ma.GLEET_Actor().cuda()
for gen in range(30):
    x = eng.feature.clone()
    bad = ~torch.isfinite(x)
    if bad.any():
        idx = bad.nonzero()
        lanes_bad = sorted(set(idx[:, 0].tolist())); feats = sorted(set(idx[:, 2].tolist()))
        print('gen', gen, 'non-finite state: lanes', lanes_bad[:10], 'fids', sorted({probs[lanes[l]].fid for l in lanes_bad}), 'feature cols', feats, 'example', x[idx[0,0], idx[0,1]].tolist())
        break
    with torch.no_grad():
        mu, sigma, lg = actor.distribution(x)
    if not torch.isfinite(mu).all():
        b = (~torch.isfinite(mu)).nonzero()[0]
        print('gen', gen, 'non-finite mu at', b.tolist(), 'state row', x[b[0], b[1]].tolist(), 'absmax state', float(x.abs().max()))
        break
    a = actor(x)
    eng.update(a)
else:
    print('30 generations: states and mu finite; absmax state', float(eng.feature.abs().max()))
