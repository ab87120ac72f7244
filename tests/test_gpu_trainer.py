"""GPU: sharded LDE training rollouts (metabox_b200/trainer.py): trajectories collected on the GPU, REINFORCE update."""
import argparse

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_lde_train_steps_collect_on_gpu_and_update_the_policy():
    import torch
    from metabox_b200 import problem as mp, trainer as mt
    cfg = argparse.Namespace(dim=10, maxFEs=20000, n_logpoint=50, log_interval=400)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob-noisy', 10, 5)
    torch.manual_seed(0)
    T = mt.LDE_Trainer(cfg, tr.data, NP=50, seed=3)
    before = [p.detach().clone() for p in T.net.parameters()]
    traj = T.collect([0, 1, 2])
    assert traj['inputs'].shape == (50, 60, 60) and traj['actions'].shape == (50, 60, 100) and traj['rewards'].shape == (50, 60)
    assert torch.all((traj['actions'] >= 0) & (traj['actions'] <= 1)) and torch.all(torch.isfinite(traj['rewards']))
    assert torch.all(traj['engine'].fes == 50 + 50 * 50)                      # init + 50 generations of NP = 50
    # hidden state is carried: h at step t+1 is the cell output of step t
    mu, sg, h1, c1 = T.net.lstm_cell(traj['inputs'][0], traj['hs'][0], traj['cs'][0])
    torch.testing.assert_close(h1, traj['hs'][1], rtol=1e-5, atol=1e-6)
    infos = [T.train_step([0, 1, 2, 3]), T.train_step([4, 5])]
    assert infos[-1]['learn_steps'] == 2 and infos[0]['problems'] == 4 and all(np.isfinite(i['loss']) for i in infos)
    assert len(infos[0]['gbest']) == 4 and all(g <= n for g, n in zip(infos[0]['gbest'], infos[0]['normalizer']))
    changed = [float((a - b.detach()).abs().max()) for a, b in zip(before, T.net.parameters())]
    assert max(changed) > 1e-4 and all(np.isfinite(c) for c in changed)
    # problems_per_step = 1 is the reference's update rule; the loss it optimises is the reference formula (CPU test)
    assert T.train_step([7])['problems'] == 1
