"""GPU: the drop-in path end to end.  The reference-API Trainer over the CUDA env + agent classes
and the OracleTrainer over the CPU oracle run the same episodes (same config, same seed, same
initial weights from the global NumPy stream, same action uniforms) and must stay in lock-step:
identical actions, rewards, dones; pi/v/returns within 1e-5; weights after every update."""
import numpy as np
import pytest
import torch

from helpers import CFG, load_cfg
from oracle.cacc import OracleCACC
from oracle.trainer import Counter as OCounter, OracleAgent, OracleTrainer

pytestmark = pytest.mark.gpu


class Stream:
    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def __call__(self):
        return self.rs.random_sample()


class Rec:
    """Wraps an agent to record every forward() result and backward() input."""

    def __init__(self, model):
        self.m, self.log = model, []

    def __getattr__(self, k):
        return getattr(self.m, k)

    def forward(self, *a, **k):
        out = self.m.forward(*a, **k)
        self.log.append(np.array(out, dtype=np.float64).ravel())
        return out

    def backward(self, R, *a, **k):
        self.log.append(np.asarray(R, dtype=np.float64).ravel())
        return self.m.backward(R, *a, **k)


@pytest.mark.parametrize('agent,cfg', [('ma2c_nc', None), ('ia2c', None), ('ma2c_ic3', None), ('ma2c_dial', None),
                                       ('ia2c', 'config_ia2c_slowdown.ini'),        # spatial returns, per-agent rewards
                                       ('ia2c_fp', None), ('ma2c_cu', None)])
def test_trainer_lockstep_with_oracle(agent, cfg):
    from deeprl_network_b200.agents.models import IA2C, IA2C_CU, IA2C_FP, MA2C_DIAL, MA2C_IC3, MA2C_NC
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    from deeprl_network_b200.utils import Counter, Trainer
    cls = {'ma2c_nc': MA2C_NC, 'ia2c': IA2C, 'ma2c_ic3': MA2C_IC3, 'ma2c_dial': MA2C_DIAL, 'ia2c_fp': IA2C_FP,
           'ma2c_cu': IA2C_CU}[agent]
    cp = load_cfg(cfg or CFG[agent])
    # CUDA side (weights drawn from np.random right after the env seeds it -- reference order)
    env = CACCEnv(cp['ENV_CONFIG'])
    model = cls(env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma, 1000,
                cp['MODEL_CONFIG'], seed=12)
    w0 = model.get_weights()
    tr = Trainer(env, Rec(model), Counter(10 ** 6, 10 ** 7, 10 ** 9), None, uniform_fn=Stream(99))
    # oracle side
    oenv = OracleCACC(cp['ENV_CONFIG'])
    oag = OracleAgent(agent, oenv.n_s_ls, oenv.n_a_ls, oenv.neighbor_mask, oenv.distance_mask, oenv.coop_gamma, 1000,
                      cp['MODEL_CONFIG'], seed=12)
    for k, v in oag.policy.p.items():                       # same init stream -> same weights
        np.testing.assert_array_equal(v.detach().numpy(), w0[k])
    otr = OracleTrainer(oenv, Rec(oag), OCounter(10 ** 6, 10 ** 7, 10 ** 9), uniform_fn=Stream(99))
    tr.run(max_episodes=2)
    otr.run(max_episodes=2)
    assert tr.global_counter.cur_step == otr.global_counter.cur_step
    a, b = tr.model.log, otr.model.log
    assert len(a) == len(b)
    worst = max(np.abs(x - y).max() for x, y in zip(a, b))
    assert worst < 1e-5, worst                                  # every pi, v and bootstrap R
    for da, db in zip(tr.data, otr.data):                       # logged (greedy test episode) rewards
        assert abs(da['avg_reward'] - db['avg_reward']) <= 1e-6 * abs(db['avg_reward'])
        assert da['step'] == db['step']
    w = model.get_weights()
    for k, v in oag.policy.p.items():
        np.testing.assert_allclose(w[k], v.detach().numpy(), rtol=0, atol=2e-5, err_msg=k)
    assert env.seed == oenv.seed == 12 + 4                      # quirk Q3: two resets per training episode
    if agent == 'ia2c_fp':        # the null message encoder behind FPPolicy must not have moved (layout.py docstring)
        flat = model.engine.params.cpu().numpy()
        named = np.zeros(flat.size, dtype=bool)
        for _, o, shape in model.layout.entries:
            named[o:o + int(np.prod(shape))] = True
        assert not flat[~named].any()
        assert model.engine.grads.cpu().numpy()[~named].any() == False


def test_save_load_roundtrip(tmp_path):
    from deeprl_network_b200.agents.models import MA2C_NC
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    cp = load_cfg(CFG['ma2c_nc'])
    env = CACCEnv(cp['ENV_CONFIG'])
    m = MA2C_NC(env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma, 0, cp['MODEL_CONFIG'])
    d = str(tmp_path) + '/'
    assert m.load(d) is False
    m.save(d, 120); m.save(d, 60)
    w = m.get_weights()
    m2 = MA2C_NC(env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma, 0, cp['MODEL_CONFIG'])
    assert m2.load(d) is True
    for k in w:
        np.testing.assert_array_equal(m2.get_weights()[k], w[k])
    ob = env.reset()
    p1 = m.forward(ob, True, env.get_fingerprint()); p2 = m2.forward(ob, True, env.get_fingerprint())
    np.testing.assert_array_equal(p1, p2)


def test_evaluator_csv_matches_reference(tmp_path):
    """a recorded test episode writes the reference's two CSV files column for column
    (fixtures produced by the unmodified reference env, tests/golden/make_golden.py)."""
    import pandas as pd
    from helpers import GOLDEN
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    cp = load_cfg(CFG['ma2c_nc'])
    env = CACCEnv(cp['ENV_CONFIG'])
    env.init_test_seeds([2000])
    env.train_mode = False
    env.cur_episode = 0
    out = str(tmp_path) + '/'
    env.init_data(True, False, out)
    env.reset(test_ind=0)
    acts = np.load(GOLDEN + '/eval_actions.npy')
    for t in range(len(acts)):
        _, _, d, _ = env.step(acts[t])
    assert d
    env.output_data()
    for kind in ('control', 'traffic'):
        mine = pd.read_csv(out + 'catchup_ma2c_nc_%s.csv' % kind)
        ref = pd.read_csv(GOLDEN + '/eval_catchup_ma2c_nc_%s.csv' % kind)
        assert list(mine.columns) == list(ref.columns)
        assert len(mine) == len(ref)
        for col in ref.columns:
            if not pd.api.types.is_numeric_dtype(ref[col]):
                assert (mine[col] == ref[col]).all(), col
            else:
                np.testing.assert_allclose(mine[col].values, ref[col].values, rtol=1e-9, atol=1e-9, err_msg=col)
