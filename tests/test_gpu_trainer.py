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


def test_rlepso_ppo_episodes_collect_on_gpu_and_update_actor_and_critic():
    import torch
    from metabox_b200 import problem as mp, trainer as mt
    cfg = argparse.Namespace(dim=10, maxFEs=2500, n_logpoint=50, log_interval=50)
    tr, te = mp.BBOB_Dataset.get_datasets('bbob', 10, 5)
    torch.manual_seed(0)
    T = mt.RLEPSO_Trainer(cfg, tr.data, seed=5)
    a0 = [p.detach().clone() for p in T.actor.parameters()]
    c0 = [p.detach().clone() for p in T.critic.parameters()]
    info = T.train_episodes([0, 1, 2, 3, 4, 5])
    # >= 25 generations of NP = 100 (re-initialisations charge extra FEs) -> 2 or 3 segments of 10, 3 updates each
    assert info['segments'] in (2, 3) and info['learn_steps'] == 3 * info['segments'] and not info['done']
    assert len(info['gbest']) == 6 and all(g <= n for g, n in zip(info['gbest'], info['normalizer']))
    assert all(np.isfinite(r) for r in info['return']) and np.isfinite(info['reinforce_loss']) and np.isfinite(info['baseline_loss'])
    da = max(float((a - b.detach()).abs().max()) for a, b in zip(a0, T.actor.parameters()))
    dc = max(float((a - b.detach()).abs().max()) for a, b in zip(c0, T.critic.parameters()))
    assert 1e-6 < da < 1e-3 and 1e-6 < dc < 1e-3                  # Adam, lr 1e-5, <= 9 steps
    # one segment: shapes, live mask, returns are sums of +-1 rewards
    from metabox_b200.optimizer import RLEPSO_Engine
    eng = RLEPSO_Engine(T.pack, np.arange(4, dtype=np.int32), 2500, 50, 50, seed=1)
    eng.init_population()
    seg = T.collect_segment(eng)
    assert seg['states'].shape == (10, 4, 1) and seg['actions'].shape == (10, 4, 35) and seg['logp'].shape == (10, 4)
    assert torch.all(seg['weight'] == 1) and torch.all(seg['rewards'].abs() == 1)
    assert torch.all((seg['actions'] >= 0) & (seg['actions'] <= 1))
    torch.testing.assert_close(seg['states'][1:, :, 0], seg['states'][1:, :, 0].clamp(0, 1))
    # the learning budget stops the episode like the reference (:267-271)
    cfg2 = argparse.Namespace(dim=10, maxFEs=2500, n_logpoint=50, log_interval=50, max_learning_step=T.learning_time + 4)
    T.config = cfg2
    info2 = T.train_episodes([6, 7])
    assert info2['done'] and info2['learn_steps'] == cfg2.max_learning_step
