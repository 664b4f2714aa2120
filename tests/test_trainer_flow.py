"""CPU: the rollout control flow pinned to the UNMODIFIED reference Trainer.
tests/golden/make_golden.py ran the reference's own `utils.Trainer` + `Counter` + `CACCEnv` around a scripted agent
(helpers.ScriptedAgent) and stored the trace of every agent call with all arguments.  Here the oracle trainer and
the product `Trainer` (the drop-in mirror, driven over the CPU oracle env -- it only needs the env API) must
reproduce that trace bit for bit: same observations, done flags, fingerprints, sampled actions (np.random stream),
rewards, values, bootstrap targets, episode boundaries, logged test-episode rewards and final env seed."""
import numpy as np
import pytest

from helpers import ScriptedAgent, golden, load_cfg
from oracle.cacc import OracleCACC
from oracle.trainer import Counter as OracleCounter, OracleTrainer

CASES = ['trainer_ma2c_nc_catchup', 'trainer_ia2c_slowdown', 'trainer_ia2c_fp_catchup']


def _setup(name):
    g = golden(name)
    cp = load_cfg(str(g['ini']))
    env = OracleCACC(cp['ENV_CONFIG'])
    agent = ScriptedAgent(env.agent, env.n_agent, env.n_a, cp['MODEL_CONFIG'].getint('batch_size'))
    return g, env, agent


def _check(g, env, agent, data, cur_step):
    trace = np.array(agent.trace)
    assert trace.shape == g['trace'].shape
    np.testing.assert_array_equal(trace, g['trace'])
    got = np.array([[d['step'], d['avg_reward'], d['std_reward']] for d in data])
    np.testing.assert_array_equal(got, g['data'])
    assert env.seed == int(g['seed_after']) and cur_step == int(g['cur_step'])


@pytest.mark.parametrize('name', CASES)
def test_oracle_trainer_reproduces_reference_trainer(name):
    g, env, agent = _setup(name)
    counter = OracleCounter(int(g['total_step']), 10 ** 9, 10 ** 9)
    tr = OracleTrainer(env, agent, counter)
    tr.run()
    _check(g, env, agent, tr.data, counter.cur_step)


@pytest.mark.parametrize('name', CASES)
def test_product_trainer_reproduces_reference_trainer(name, tmp_path):
    from deeprl_network_b200.utils import Counter, Trainer
    g, env, agent = _setup(name)
    counter = Counter(int(g['total_step']), 10 ** 9, 10 ** 9)
    tr = Trainer(env, agent, counter, None, output_path=str(tmp_path) + '/')
    tr.run()
    _check(g, env, agent, tr.data, counter.cur_step)
