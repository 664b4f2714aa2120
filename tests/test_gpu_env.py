"""GPU: K1 (cacc_reset / cacc_step kernels through the C ABI and the CACCEnv mirror) against the
committed reference trajectories and the oracle.  The reference env is float64: state and rewards
must agree to float64 round-off (cos() may differ by an ulp), observations as float32."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_cfg
from oracle.cacc import OracleCACC

pytestmark = pytest.mark.gpu
FILES = sorted(glob.glob(os.path.join(GOLDEN, 'env_*.npz')))


@pytest.mark.parametrize('path', FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_env_api_matches_reference_trajectory(path):
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    g = np.load(path, allow_pickle=True)
    cp = load_cfg(str(g['ini']), **eval(str(g['over'])))
    env = CACCEnv(cp['ENV_CONFIG'])
    for ep in range(int(g['n_ep'])):
        if bool(g['test_mode']):
            env.train_mode = True; env.reset(); env.train_mode = False
            ob = env.reset(test_ind=-1)
        else:
            ob = env.reset()
        assert env.seed == int(g['ep%d_seed_after' % ep])
        np.testing.assert_array_equal(env.hs[:, 0].cpu().numpy(), g['ep%d_h0' % ep])
        np.testing.assert_array_equal(env.vs[:, 0].cpu().numpy(), g['ep%d_v0' % ep])
        ref_obs = g['ep%d_obs' % ep]
        np.testing.assert_allclose(np.concatenate(ob), ref_obs[0].astype(np.float32), rtol=0, atol=1e-6)
        acts = g['ep%d_acts' % ep]
        n_exact = 0
        for t in range(len(acts)):
            ob, r, d, gr = env.step(acts[t])
            assert d == bool(g['ep%d_done' % ep][t]), t
            ref_g = g['ep%d_greward' % ep][t]
            assert abs(gr - ref_g) <= 1e-9 * max(1.0, abs(ref_g)), (t, gr, ref_g)
            np.testing.assert_allclose(np.broadcast_to(r, (env.n_agent,)), g['ep%d_rew' % ep][t], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(np.concatenate(ob), ref_obs[t + 1].astype(np.float32), rtol=0, atol=1e-6)
            n_exact += int(gr == ref_g)
        np.testing.assert_allclose(env.hs[:, 0].cpu().numpy(), g['ep%d_hs' % ep][-1], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(env.vs[:, 0].cpu().numpy(), g['ep%d_vs' % ep][-1], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(env.us[:, 0].cpu().numpy(), g['ep%d_us' % ep][-1], rtol=1e-9, atol=1e-9)
        # CUDA's and glibc's float64 cos() differ by an ulp now and then, so only part of the rewards are
        # bit-identical; everything stays within float64 round-off of the reference (asserts above)
        assert n_exact > 0


@pytest.mark.parametrize('ini', ['config_ma2c_nc_catchup.ini', 'config_ma2c_cnet_slowdown.ini', 'config_ia2c_slowdown.ini'])
def test_batched_envs_match_per_env_oracle(ini):
    """B envs with different initial uniforms and different action streams == B independent oracles;
    per-env done / collision latches and per-env time."""
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    B, steps = 37, 310
    cp = load_cfg(ini)
    env = CACCEnv(cp['ENV_CONFIG'], n_env=B)
    rs = np.random.RandomState(5)
    u = rs.rand(1, B)
    acts = rs.randint(0, 4, size=(steps, env.n_agent, B)).astype(np.int32)
    acts[:, :, 0] = 3                               # env 0 never collides
    acts[:, :, 1] = 0
    acts[:, :, 2] = (np.arange(steps)[:, None] + np.arange(env.n_agent)[None, :]) % 4     # collides (golden 'cyc')
    env.reset_device(u01=torch.as_tensor(u).to(env.device))
    oracles = []
    for b in range(B):
        o = OracleCACC(cp['ENV_CONFIG']); o.reset(u01=u[0, b]); oracles.append(o)
    np.testing.assert_array_equal(env.hs.cpu().numpy(), np.stack([o.hs_cur for o in oracles], 1))
    alive = np.ones(B, bool)
    for t in range(steps):
        env.step_device(torch.as_tensor(acts[t]).to(env.device))
        obs = env.obs_dev[..., :5].cpu().numpy()
        rew = env.reward_dev.cpu().numpy(); grew = env.greward_dev.cpu().numpy(); done = env.done_dev.cpu().numpy()
        for b in range(B):
            if not alive[b]:
                continue
            ob, r, d, gr = oracles[b].step(acts[t, :, b])
            base = np.stack([x[:5] for x in ob])
            np.testing.assert_allclose(obs[:, b], base.astype(np.float32), rtol=0, atol=1e-6)
            assert abs(grew[b] - gr) <= 1e-9 * max(1, abs(gr))
            np.testing.assert_allclose(rew[:, b], np.broadcast_to(r, rew[:, b].shape), rtol=1e-9, atol=1e-9)
            assert bool(done[b]) == d
            alive[b] = not d
    assert (~alive).sum() >= 1 and not alive[2] and alive[0]     # the 'cyc' env collided and ended, env 0 did not
    tt = env.t_dev.cpu().numpy()
    assert tt[0] == steps


def test_masked_reset_and_philox_reset():
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    cp = load_cfg('config_ma2c_nc_catchup.ini')
    env = CACCEnv(cp['ENV_CONFIG'], n_env=64)
    env.reset_device(u01=None, philox_seed=7)
    h0 = env.hs[0].cpu().numpy()
    assert np.all((h0 >= 30.0) & (h0 < 50.0)) and len(np.unique(h0)) == 64      # h*(1.5+U), all different
    assert np.all(env.hs[1:].cpu().numpy() == 20.0) and np.all(env.fp_dev.cpu().numpy() == 0.25)
    a = torch.full((8, 64), 3, dtype=torch.int32, device=env.device)
    for _ in range(5):
        env.step_device(a)
    mask = torch.zeros(64, device=env.device); mask[::2] = 1
    before = env.hs.clone()
    env.reset_device(u01=None, mask=mask, philox_seed=7)
    t = env.t_dev.cpu().numpy()
    assert np.all(t[::2] == 0) and np.all(t[1::2] == 5)
    assert torch.equal(env.hs[:, 1::2], before[:, 1::2])
    h1 = env.hs[0, ::2].cpu().numpy()
    assert np.all(h1 != h0[::2])                     # next episode of the same env draws a new uniform
    ep = env.episode_dev.cpu().numpy()
    assert np.all(ep[::2] == 2) and np.all(ep[1::2] == 1)


def test_multi_platoon_stub():
    """cfg5 dynamics stub: 25 vehicles = 5 independent platoons of 5 == five 5-vehicle oracles."""
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    cp = load_cfg('config_ma2c_nc_grid5x5_stub.ini', n_env=3)
    env = CACCEnv(cp['ENV_CONFIG'])
    assert env.n_agent == 25 and env.neighbor_mask.sum() == 80
    rs = np.random.RandomState(0)
    u = rs.rand(5, 3)
    env.reset_device(u01=torch.as_tensor(u).to(env.device))
    cp1 = load_cfg('config_ma2c_nc_grid5x5_stub.ini', n_vehicle=5)
    acts = rs.randint(0, 4, size=(40, 25, 3)).astype(np.int32)
    orc = [[OracleCACC(cp1['ENV_CONFIG']) for _ in range(5)] for _ in range(3)]
    for b in range(3):
        for p in range(5):
            orc[b][p].reset(u01=u[p, b])
    for t in range(40):
        env.step_device(torch.as_tensor(acts[t]).to(env.device))
        hs = env.hs.cpu().numpy()
        for b in range(3):
            if env.collision_dev[b].item():
                continue
            for p in range(5):
                orc[b][p].step(acts[t, 5 * p:5 * p + 5, b])
                np.testing.assert_allclose(hs[5 * p:5 * p + 5, b], orc[b][p].hs_cur, rtol=1e-12)
