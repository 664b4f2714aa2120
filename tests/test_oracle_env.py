"""CPU: the oracle env restatement is pinned bit-for-bit to the unmodified reference
(fixtures produced by tests/golden/make_golden.py) and to the SURVEY 8(c) known answers."""
import glob
import os

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, load_cfg
from oracle.cacc import OracleCACC, leader_speed, np_pairwise_sum

FILES = sorted(glob.glob(os.path.join(GOLDEN, 'env_*.npz')))
# ia2c_fp / ma2c_cu trajectories with the fingerprints the trainer would set before every step
FP_FILES = sorted(glob.glob(os.path.join(GOLDEN, 'envfp_*.npz')))


def _run(g):
    over = eval(str(g['over']))
    cp = load_cfg(str(g['ini']), **over)
    env = OracleCACC(cp['ENV_CONFIG'])
    for ep in range(int(g['n_ep'])):
        if bool(g['test_mode']):
            env.train_mode = True; env.reset(); env.train_mode = False
            ob = env.reset(test_ind=-1)
        else:
            ob = env.reset()
        assert env.seed == int(g['ep%d_seed_after' % ep])
        np.testing.assert_array_equal(env.hs_cur, g['ep%d_h0' % ep])
        np.testing.assert_array_equal(env.vs_cur, g['ep%d_v0' % ep])
        obs = [np.concatenate(ob)]
        acts = g['ep%d_acts' % ep]
        fps = g['ep%d_fps' % ep] if ('ep%d_fps' % ep) in g.files else None
        for t in range(len(acts)):
            if fps is not None:
                env.update_fingerprint(fps[t])
            ob, r, d, gr = env.step(acts[t])
            obs.append(np.concatenate(ob))
            assert gr == g['ep%d_greward' % ep][t]
            assert d == bool(g['ep%d_done' % ep][t])
            np.testing.assert_array_equal(np.broadcast_to(r, (env.n_agent,)), g['ep%d_rew' % ep][t])
            np.testing.assert_array_equal(env.hs_cur, g['ep%d_hs' % ep][t])
        np.testing.assert_array_equal(np.array(obs), g['ep%d_obs' % ep])
        yield env, g, ep


@pytest.mark.parametrize('path', FILES + FP_FILES, ids=[os.path.basename(f)[:-4] for f in FILES + FP_FILES])
def test_oracle_matches_reference_trajectory(path):
    assert len(FILES) >= 10 and len(FP_FILES) == 3
    g = np.load(path, allow_pickle=True)
    for _ in _run(g):
        pass


def test_known_answers_survey_8c():
    g = np.load(os.path.join(GOLDEN, 'env_nc_catchup_const3.npz'), allow_pickle=True)
    assert abs(g['ep0_h0'][0] - 33.08325685) < 1e-7
    assert abs(g['ep0_greward'][0] - (-171.5323408189)) < 1e-9
    assert abs(g['ep0_greward'].sum() - (-17787.5552013954)) < 1e-7
    g = np.load(os.path.join(GOLDEN, 'env_nc_catchup_cyc.npz'), allow_pickle=True)
    assert len(g['ep0_done']) == 240 and abs(g['ep0_greward'].sum() - (-230525.1115080572)) < 1e-6
    g = np.load(os.path.join(GOLDEN, 'env_ic3_slowdown_const3.npz'), allow_pickle=True)
    assert abs(g['ep0_v0'][0] - 24.812442635695085) < 1e-12
    assert abs(g['ep0_greward'].sum() - (-104209.3542998925)) < 1e-6
    g = np.load(os.path.join(GOLDEN, 'env_ic3_slowdown_const0.npz'), allow_pickle=True)
    assert len(g['ep0_done']) == 120 and g['ep0_greward'][-1] == -8000.0


def test_leader_profile_closed_form():
    """The kernel's analytic v0s[t] equals np.linspace-based profile of the reference, bit for bit."""
    g = np.load(os.path.join(GOLDEN, 'env_ic3_slowdown_const3.npz'), allow_pickle=True)
    v0s, v_init = g['ep0_v0s'], g['ep0_v0'][0]
    for t in range(len(v0s)):
        assert leader_speed('slowdown', v_init, 15.0, t) == v0s[t], t
    g = np.load(os.path.join(GOLDEN, 'env_nc_catchup_const3.npz'), allow_pickle=True)
    assert all(leader_speed('catchup', 15.0, 15.0, t) == x for t, x in enumerate(g['ep0_v0s']))


def test_np_sum_order_restated():
    rs = np.random.RandomState(1)
    for n in (1, 5, 7, 8, 9, 16, 25, 31):
        x = rs.randn(n) * 1e3
        assert np_pairwise_sum(x) == np.sum(x), n


def test_ia2c_observation_is_own_plus_neighbours():
    cp = load_cfg('config_ia2c_catchup.ini')
    env = OracleCACC(cp['ENV_CONFIG'])
    ob = env.reset()
    assert [len(o) for o in ob] == [10, 15, 15, 15, 15, 15, 15, 10] == env.n_s_ls
    cp2 = load_cfg('config_ma2c_nc_catchup.ini')
    env2 = OracleCACC(cp2['ENV_CONFIG'])
    ob2 = env2.reset()
    np.testing.assert_array_equal(ob[3], np.concatenate([ob2[3], ob2[2], ob2[4]]))


def test_fingerprints_sit_at_the_end_of_the_ia2c_fp_observation():
    """envs/cacc_env.py:74-77 on the reference's own trajectory: [own 5 | neighbours 5 each | fingerprints 4 each]."""
    g = np.load(os.path.join(GOLDEN, 'envfp_ia2c_fp_catchup_rand.npz'), allow_pickle=True)
    obs, fps = g['ep0_obs'], g['ep0_fps']
    widths = [5 * 2 + 4, 5 * 3 + 8, 5 * 3 + 8, 5 * 3 + 8, 5 * 3 + 8, 5 * 3 + 8, 5 * 3 + 8, 5 * 2 + 4]
    assert obs.shape[1] == sum(widths)
    np.testing.assert_array_equal(obs[0][10:14], np.full(4, 0.25))            # reset: uniform fingerprint of agent 1
    o3 = obs[5][sum(widths[:3]):sum(widths[:4])]                               # agent 3 after 5 steps: neighbours 2 and 4
    np.testing.assert_array_equal(o3[15:19], fps[4][2])
    np.testing.assert_array_equal(o3[19:23], fps[4][4])
