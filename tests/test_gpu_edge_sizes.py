"""GPU: population sizes at the edges of the supported range (MBX_MAX_NP = 128, minimum 5 / 6, a ragged 37 that is neither a
multiple of 4, 5 nor 32), dim 30: every generation is replayed by the oracle (Philox twin) from the GPU's previous state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FIDS = [1, 7, 15, 21]


def _setup(dim=30):
    from metabox_b200 import problem as mp
    from oracle import bbob as ob
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', dim, 5)
    probs = {p.fid: p for p in tr.data + te.data}
    otr, ote = ob.make_suite('bbob', dim, 5.0)
    oin = {i.fid: i for i in otr + ote}
    return mp.InstancePack([probs[f] for f in FIDS]), [oin[f] for f in FIDS]


@pytest.mark.parametrize('NP', [5, 35, 125])          # RLEPSO needs a multiple of its 5 coefficient groups (rlepso_optimizer.py:27)
def test_rlepso_population_sizes(NP):
    import torch
    from metabox_b200 import optimizer as mo
    from oracle import bbob as ob, rlepso as orl
    from test_gpu_rlepso import _check
    pack, oins = _setup()
    seed, dim = 99, 30
    eng = mo.RLEPSO_Engine(pack, list(range(len(FIDS))), 60000, 50, seed=seed, NP=NP)
    cfg = orl.Config(dim=dim, maxFEs=60000, NP=NP)
    eng.init_population()
    g = np.random.RandomState(NP)
    for gen in range(8):
        action = g.rand(len(FIDS), 35).astype(np.float32)
        if gen == 4:
            eng.pni += 40                                   # provoke re-initialisation
        prev = [eng.read_state(b) for b in range(len(FIDS))]
        dbg = torch.full((len(FIDS), NP, dim), -1, dtype=torch.int32, device=eng.device)
        eng.update(torch.as_tensor(action).to(eng.device), dbg=dbg)
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = orl.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], pni=p['pni'],
                          gbest_x=p['gbest_x'], gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']), no_improve=0,
                          fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            aux = {}
            N, _, _, _ = orl.update(cfg, S, action[b], ev, I.optimum, -5., 5., orl.PhiloxTape(NP, dim, seed, b, gen), aux)
            np.testing.assert_array_equal(dbg[b].cpu().numpy(), aux['fdr_target'], err_msg=f'NP {NP} gen {gen} lane {b} FDR targets')
            _check(eng.read_state(b), N, f'NP {NP} gen {gen} lane {b}')


def test_rlepso_rejects_sizes_it_cannot_group():
    import torch
    from metabox_b200 import _lib, optimizer as mo
    pack, _ = _setup()
    for NP in (37, 130):
        with pytest.raises(_lib.MbxError):
            eng = mo.RLEPSO_Engine(pack, [0], 60000, 50, NP=NP)
            eng.init_population()
            eng.update(torch.rand(1, 35).to(eng.device))


@pytest.mark.parametrize('NP', [5, 37, 128])
def test_gleet_population_sizes(NP):
    import torch
    from metabox_b200 import optimizer as mo
    from oracle import bbob as ob, gleet as og
    from test_gpu_gleet import _check
    pack, oins = _setup()
    seed, dim = 12, 30
    eng = mo.GLEET_Engine(pack, list(range(len(FIDS))), 60000, 50, seed=seed, NP=NP)
    cfg = og.Config(dim=dim, maxFEs=60000, NP=NP)
    eng.init_population()
    g = np.random.RandomState(NP)
    for gen in range(8):
        action = g.rand(len(FIDS), NP).astype(np.float32)
        prev = [eng.read_state(b) for b in range(len(FIDS))]
        feat, reward, done = eng.update(torch.as_tensor(action).to(eng.device))
        feat = feat.cpu().numpy()
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = og.State(x=p['x'], v=p['v'], pbest_x=p['pbest_x'], c_cost=p['c_cost'], pbest=p['pbest'], pni=p['pni'],
                         gbest_x=p['gbest_x'], gbest_val=np.float32(p['gbest_val']), gbest_idx=int(p['gbest_idx']),
                         no_improve=int(p['no_improve']), w=float(p['w']), max_cost=np.float32(p['max_cost']),
                         pbest_feature=p['pbest_feature'], gbest_feature=p['gbest_feature'], fes=int(p['fes']),
                         log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, state, r, d = og.update(cfg, S, action[b], ev, I.optimum, -5., 5., og.PhiloxTape(NP, dim, seed, b, gen))
            _check(eng.read_state(b), feat[b], N, state, f'NP {NP} gen {gen} lane {b}')


@pytest.mark.parametrize('NP', [6, 37, 128])
def test_lde_population_sizes(NP):
    import torch
    from metabox_b200 import optimizer as mo
    from oracle import bbob as ob, lde as ol
    from test_gpu_lde import _check
    pack, oins = _setup()
    seed, dim = 31, 30
    eng = mo.LDE_Engine(pack, list(range(len(FIDS))), 60000, 50, seed=seed, NP=NP)
    cfg = ol.Config(dim=dim, maxFEs=60000, NP=NP)
    eng.init_population()
    g = np.random.RandomState(NP)
    for gen in range(8):
        action = g.rand(len(FIDS), 2 * NP).astype(np.float32)
        prev = [eng.read_state(b) for b in range(len(FIDS))]
        fg, rg, dg = eng.update(torch.as_tensor(action).to(eng.device))
        fg = fg.cpu().numpy()
        for b, I in enumerate(oins):
            ev = lambda x, draws, I=I: ob.func_true(I, x)
            p = prev[b]
            S = ol.State(pop=p['pop'], fit=p['fit'], hist_sum=p['hist_sum'], hist_cnt=int(p['hist_cnt']), gbest=p['gbest'],
                         fes=int(p['fes']), log_index=int(p['log_index']), curve=list(p['curve'][:int(p['curve_len'])]))
            N, feat, reward, done = ol.update(cfg, S, action[b], ev, I.optimum, -5., 5., ol.PhiloxTape(NP, dim, seed, b, gen, 0))
            _check(eng.read_state(b), N, fg[b], feat, f'NP {NP} gen {gen} lane {b}')
