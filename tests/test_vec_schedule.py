"""Host logic of the batched loop (no GPU): the learning-rate schedule counts ENVIRONMENT steps, so a linear
`lr_decay` reaches lr_min when n_step * n_env * world * updates == total_step, and updates must tile
the episode (the per-env auto-reset only looks at the last step of an update)."""
import types

import pytest
import torch

from deeprl_network_b200.agents.utils import Scheduler
from deeprl_network_b200.utils import VecTrainer


class _Loop(VecTrainer):
    def _one_update(self, uniforms=None):
        pass


def _stubs(n_env=4, world=2, n_step=60, total=60 * 4 * 2 * 25, T=600, env_batch=60):
    env = types.SimpleNamespace(n_env=n_env, T=T, batch_size=env_batch, agent='ma2c_nc', train_mode=False)
    engine = types.SimpleNamespace(world=world, lr_dev=torch.zeros(1))
    model = types.SimpleNamespace(n_env=n_env, n_step=n_step, engine=engine,
                                  lr_scheduler=Scheduler(5e-4, 1e-4, total, decay='linear'))
    return env, model, total


def test_linear_lr_reaches_lr_min_at_total_step():
    env, model, total = _stubs()
    loop = _Loop(env, model, graph=False)
    lrs, steps = [], 0
    while steps < total:
        loop.update()
        steps += model.n_step * env.n_env * model.engine.world
        lrs.append(float(model.engine.lr_dev.item()))
    assert lrs[0] == pytest.approx(5e-4 * (1 - 480 / total), rel=1e-6)
    assert lrs[len(lrs) // 2] == pytest.approx(5e-4 * (1 - (len(lrs) // 2 + 1) * 480 / total), rel=1e-6)
    assert lrs[-1] == pytest.approx(1e-4, rel=1e-6) and min(lrs) >= 1e-4 * (1 - 1e-6)      # lr_dev is fp32
    assert all(a >= b for a, b in zip(lrs, lrs[1:]))


def test_updates_must_tile_the_episode():
    env, model, _ = _stubs(n_step=45)
    with pytest.raises(AssertionError):
        _Loop(env, model, graph=False)
    env, model, _ = _stubs(n_step=30, env_batch=60)        # done can only fire at multiples of 60: fine
    _Loop(env, model, graph=False)
