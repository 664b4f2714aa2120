"""GPU: the drop-in path -- CUDA env, agent
classes, Trainer -- directly against the traces of the UNMODIFIED reference executed on the TF shim
(tests/golden/tfnet_*.npz, see tests/test_tfnet_parity.py), without the oracle in between: same initial weights
from the same NumPy stream, every pi / v / bootstrap R within 1e-5, same logged rewards, trained weights within 2e-5."""
import hashlib
import os

import numpy as np
import pytest

from helpers import golden, load_cfg

pytestmark = pytest.mark.gpu

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
def test_drop_in_path_follows_reference_on_tf_shim(name):
    import main
    from deeprl_network_b200.envs.cacc_env import CACCEnv
    from deeprl_network_b200.utils import Counter, Trainer
    g = golden(name)
    cp = load_cfg(str(g['ini']))
    env = CACCEnv(cp['ENV_CONFIG'])
    model = main.init_agent(env, cp['MODEL_CONFIG'], 10 ** 6, 12)
    w0 = model.get_weights()
    for n in (str(x) for x in g['names']):
        assert hashlib.sha256(np.ascontiguousarray(w0[n]).tobytes()).hexdigest() == str(g['w0sha/' + n]), n
    rec = Rec(model)
    counter = Counter(int(g['total_step']), 10 ** 9, 10 ** 9)
    tr = Trainer(env, rec, counter, None)
    tr.run()
    assert counter.cur_step == int(g['cur_step']) and env.seed == int(g['seed_after'])
    trace = np.concatenate(rec.log)
    assert trace.shape == g['trace'].shape
    assert np.abs(trace - g['trace']).max() < 1e-5
    got = np.array([[d['step'], d['avg_reward'], d['std_reward']] for d in tr.data])
    np.testing.assert_allclose(got, g['data'], rtol=1e-6)
    w1 = model.get_weights()
    assert max(np.abs(w1[str(n)] - g['w1/' + str(n)]).max() for n in g['names']) < 2e-5
