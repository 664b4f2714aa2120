"""CPU: the oracle's network restatement (oracle/nets.py: encoders, message passing, LSTM cell, heads, loss, autodiff,
global-norm clip, RMSProp) against the UNMODIFIED reference network code.  TensorFlow cannot be installed, so
tests/golden/make_golden.py executed the reference's own env + Trainer + agent + policy + layer source with `tensorflow`
replaced by tests/golden/tf_shim.py (the ~35 TF primitives restated on PyTorch-CPU) and stored the initial weights,
every pi / v / bootstrap R the Trainer saw during two training episodes (with their interleaved greedy test
episodes), and the weights after the 8 updates.  The oracle trainer must follow that run:
  * same initial weights from the same NumPy stream (variable creation order, orthogonal init) -- exact;
  * same sampled actions / episode lengths / logged rewards (so pi agrees to within the sampling margins) -- exact;
  * pi, v, R within 1e-5; weights after training within 2e-5 (same fp32 arithmetic, different op order)."""
import hashlib

import numpy as np
import pytest

from helpers import golden, load_cfg
from oracle.cacc import OracleCACC
from oracle.trainer import Counter, OracleAgent, OracleTrainer

CASES = ['tfnet_ma2c_nc_catchup', 'tfnet_ia2c_slowdown', 'tfnet_ia2c_fp_catchup', 'tfnet_ma2c_ic3_slowdown',
         'tfnet_ma2c_dial_catchup', 'tfnet_ma2c_cu_catchup']


class Rec:
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


@pytest.mark.parametrize('name', CASES)
def test_oracle_follows_reference_networks_on_tf_shim(name):
    g = golden(name)
    cp = load_cfg(str(g['ini']))
    env = OracleCACC(cp['ENV_CONFIG'])
    agent = OracleAgent(env.agent, env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma,
                        10 ** 6, cp['MODEL_CONFIG'], seed=12)
    names = [str(n) for n in g['names']]
    assert sorted(names) == sorted(agent.policy.names)                    # reference variable names, one for one
    for n in names:                                                       # creation order + ortho init reproduce w0
        w = np.ascontiguousarray(agent.policy.p[n].detach().numpy())
        assert hashlib.sha256(w.tobytes()).hexdigest() == str(g['w0sha/' + n]), n
    w0 = {n: agent.policy.p[n].detach().numpy().copy() for n in names}
    rec = Rec(agent)
    counter = Counter(int(g['total_step']), 10 ** 9, 10 ** 9)
    tr = OracleTrainer(env, rec, counter)
    tr.run()
    assert counter.cur_step == int(g['cur_step']) and env.seed == int(g['seed_after'])
    trace = np.concatenate(rec.log)
    assert trace.shape == g['trace'].shape
    assert np.abs(trace - g['trace']).max() < 1e-5
    got = np.array([[d['step'], d['avg_reward'], d['std_reward']] for d in tr.data])
    np.testing.assert_allclose(got, g['data'], rtol=1e-9)
    worst = max(np.abs(agent.policy.p[n].detach().numpy() - g['w1/' + n]).max() for n in names)
    moved = max(np.abs(g['w1/' + n] - w0[n]).max() for n in names)
    print('%s: max |pi,v,R| deviation %.2e, max weight deviation after training %.2e (weights moved by %.2e)' % (
        name, np.abs(trace - g['trace']).max(), worst, moved))
    assert worst < 2e-5 and moved > 1e-3, (worst, moved)
