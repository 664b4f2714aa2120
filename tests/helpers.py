"""Shared fixtures for the parity tests (oracle <-> CUDA path)."""
import configparser
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CFG = {'ma2c_nc': 'config_ma2c_nc_catchup.ini', 'ia2c': 'config_ia2c_catchup.ini',
       'ma2c_ic3': 'config_ma2c_cnet_slowdown.ini', 'ma2c_dial': 'config_ma2c_dial_catchup.ini',
       'ia2c_fp': 'config_ia2c_fp_slowdown.ini', 'ma2c_cu': 'config_ia2c_cu_catchup.ini'}


def load_cfg(name, **env_over):
    cp = configparser.ConfigParser()
    assert cp.read(os.path.join(ROOT, 'config', name)), name
    for k, v in env_over.items():
        cp['ENV_CONFIG'][k] = str(v)
    return cp


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=True)


def random_params(layout_or_shapes, seed=0, scale=0.3):
    """Generic (non-orthogonal, non-zero-bias) weights so every term of the graph is exercised."""
    rs = np.random.RandomState(seed)
    shapes = layout_or_shapes
    return {n: (rs.standard_normal(s) * scale / np.sqrt(max(s[0], 1) if len(s) == 2 else 1.0)).astype(np.float32)
            for n, s in shapes}


class ScriptedAgent:
    """A deterministic stand-in for the IA2C / MA2C agent classes, used to pin the TRAINER control flow: it has
    the agent API the reference Trainer calls (forward / add_transition / backward / reset, `n_step`, `n_agent`,
    `sess`), answers every call with a closed-form function of its inputs and of how many policy calls it has
    seen since reset() (a stand-in for the recurrent state, so that quirks Q1/Q2 are visible), and appends every
    call with all its arguments to one flat float64 trace."""

    class _Sess:
        def run(self, *a, **k):
            return None

    def __init__(self, name, n_agent, n_a, n_step):
        self.name, self.n_agent, self.n_a, self.n_step = name, n_agent, n_a, n_step
        self.sess = self._Sess()
        self.k = 0                       # policy calls since reset()
        self.trace = []
        rs = np.random.RandomState(1234)
        self.w = rs.randn(n_agent, 5, n_a)

    def _rec(self, code, *parts):
        self.trace.append(float(code))
        for p in parts:
            self.trace.extend(np.asarray(p, dtype=np.float64).ravel().tolist())

    def reset(self):
        self._rec(1)
        self.k = 0

    def forward(self, ob, done, extra=None, actions=None, out_type='p'):
        if isinstance(actions, str):       # IA2C signature: forward(ob, done, nactions, 'v')
            out_type, actions = actions, None
        own = np.array([np.asarray(o, dtype=np.float64)[:5] for o in ob])
        if out_type.startswith('p'):
            self.k += 1
            z = np.einsum('if,ifa->ia', own, self.w) + 0.05 * self.k - 0.5 * float(bool(done))
            e = np.exp(z - z.max(1, keepdims=True))
            pi = e / e.sum(1, keepdims=True)
            self._rec(2, float(bool(done)), np.concatenate([np.asarray(o, dtype=np.float64) for o in ob]),
                      [] if (extra is None or self.name.startswith('ia2c')) else extra, pi)
            return [pi[i] for i in range(self.n_agent)] if self.name.startswith('ia2c') else pi
        v = own.sum(1) * 0.01 + 0.1 * self.k
        if self.name.startswith('ia2c'):
            acts = np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in extra])
        else:
            acts = np.asarray(actions, dtype=np.float64)
        self._rec(3, float(bool(done)), acts, v)
        return [v[i] for i in range(self.n_agent)] if self.name.startswith('ia2c') else v

    def add_transition(self, ob, p, action, reward, value, done):
        extra = np.concatenate([np.asarray(x, dtype=np.float64).ravel() for x in p]) if self.name.startswith('ia2c') else p
        self._rec(4, np.concatenate([np.asarray(o, dtype=np.float64) for o in ob]), extra, action,
                  np.broadcast_to(np.asarray(reward, dtype=np.float64), (self.n_agent,)), value, float(bool(done)))

    def backward(self, Rends, dt=0, summary_writer=None, global_step=None):
        self._rec(5, Rends, dt)
        return {}


# ---- scripted POLICY level (one step below ScriptedAgent): pins the agent classes' host logic ------------------
def script_pi(own, k, done, w):
    """own [n,5] float64 -> pi [n, n_a]; k = policy calls since reset (stand-in for the recurrent state)."""
    z = np.einsum('if,ifa->ia', own, w) + 0.05 * k - 0.5 * float(bool(done))
    e = np.exp(z - z.max(1, keepdims=True))
    return e / e.sum(1, keepdims=True)


def script_v(own, k):
    return own.sum(1) * 0.01 + 0.1 * k


class PolicyTrace:
    """Canonical record of what crosses the agent -> policy boundary (the TF boundary in the reference)."""

    def __init__(self, n_agent, n_a):
        self.t = []
        self.w = np.random.RandomState(4321).randn(n_agent, 5, n_a)

    def rec(self, code, *parts):
        self.t.append(float(code))
        for p in parts:
            self.t.extend(np.asarray(p, dtype=np.float64).ravel().tolist())
