"""CPU: returns/advantages, buffer protocol and LR schedule of the oracle are pinned to the
unmodified reference buffers (tests/golden/buffer_*.npz, scheduler.npz)."""
import numpy as np
import pytest

from helpers import golden
from oracle.buffers import RolloutBuffer, Scheduler, nstep_returns


@pytest.mark.parametrize('name', ['buffer_ma_global', 'buffer_ma_spatial09', 'buffer_ia_global', 'buffer_ia_spatial08'])
def test_returns_match_reference(name):
    g = golden(name)
    Rs, Advs = nstep_returns(g['r'], g['v'], g['done_post'], g['R_end'], float(g['gamma']), float(g['alpha']), g['dist'])
    np.testing.assert_array_equal(Rs, g['Rs'])
    np.testing.assert_array_equal(Advs, g['Advs'])


def test_buffer_protocol_pre_step_dones():
    g = golden('buffer_ma_global')
    buf = RolloutBuffer(float(g['gamma']), float(g['alpha']), g['dist'])
    T = len(g['r'])
    for t in range(T):
        buf.add_transition(np.zeros((8, 5)), np.zeros((8, 4)), np.zeros(8, dtype=int), g['r'][t][0], g['v'][t], bool(g['done_post'][t]))
    obs, ps, acts, dones, Rs, Advs = buf.sample_transition(g['R_end'])
    np.testing.assert_array_equal(dones, g['dones_pre'])
    assert dones[0] == False and dones[30] == True       # done after step 29 shows up before step 30
    np.testing.assert_array_equal(Rs, g['Rs'])
    assert buf.dones == [bool(g['done_post'][-1])]


def test_survey_known_answer():
    rs = np.random.RandomState(0)
    r, v = [], []
    for t in range(60):
        r.append(np.full(8, rs.randn())); rs.randn(8, 5); rs.rand(8, 4); rs.randint(0, 4, 8); v.append(rs.randn(8))
    R_end = rs.randn(8)
    dist = np.abs(np.arange(8)[:, None] - np.arange(8)[None, :])
    Rs, Advs = nstep_returns(np.array(r), np.array(v), np.zeros(60, bool), R_end, 0.99, -1, dist)
    np.testing.assert_allclose(Rs[0, :3], [-10.334589, -12.22085, -11.652565], rtol=1e-6)
    np.testing.assert_allclose(Advs[7, -2:], [-1.1333116, -1.262278], rtol=1e-6)
    assert abs(float(Rs.astype(np.float64).sum()) - (-3709.611228)) < 1e-2


def test_scheduler():
    g = golden('scheduler')
    s1 = Scheduler(5e-4, decay='constant')
    np.testing.assert_array_equal([s1.get(60) for _ in range(5)], g['const'])
    s2 = Scheduler(5e-4, 1e-4, 1e6, decay='linear')
    np.testing.assert_array_equal(np.array([s2.get(60) for _ in range(20000)][::997]), g['linear'])
    from deeprl_network_b200.agents.utils import Scheduler as S2
    s3 = S2(5e-4, 1e-4, 1e6, decay='linear')
    np.testing.assert_array_equal(np.array([s3.get(60) for _ in range(20000)][::997]), g['linear'])
