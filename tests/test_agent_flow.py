"""CPU: the agent classes' host logic (SURVEY 8a rows a9-a13 in situ: reward scaling / clipping, rollout buffers,
n-step and spatially discounted returns, lr schedule, argument marshalling) pinned to the UNMODIFIED reference.
tests/golden/make_golden.py ran the reference's own IA2C / IA2C_FP / MA2C_NC / MA2C_IC3 / MA2C_DIAL classes inside
its own Trainer, with only the TensorFlow policy objects replaced by scripted ones, and stored everything that
crossed the agent -> policy boundary (tests/golden/agent_*.npz).  The oracle agent + oracle trainer, with the same
scripted policy plugged in at the same boundary, must reproduce it bit for bit."""
import numpy as np
import pytest

from helpers import PolicyTrace, golden, load_cfg, script_pi, script_v
from oracle.cacc import OracleCACC
from oracle.trainer import Counter, OracleAgent, OracleTrainer

CASES = ['agent_ma2c_nc_catchup', 'agent_ia2c_slowdown', 'agent_ia2c_fp_slowdown', 'agent_ma2c_ic3_slowdown',
         'agent_ma2c_dial_catchup']


class ScriptedOraclePolicy:
    """OraclePolicy's interface (oracle/nets.py: forward / backward / reset with a leading env axis of 1), recording
    in the canonical layout of helpers.PolicyTrace: per agent for the IA2C family, jointly for the MA2C family."""

    def __init__(self, name, tr, nbr):
        self.name, self.tr, self.nbr = name, tr, nbr
        self.N = len(nbr)
        self.n_a = tr.w.shape[2]
        self.single = name.startswith('ia2c')
        self.k = np.zeros(self.N, dtype=int) if self.single else 0

    def reset(self):
        if self.single:
            for i in range(self.N):
                self.tr.rec(1, i)
            self.k[:] = 0
        else:
            self.tr.rec(1)
            self.k = 0

    def forward(self, obs, done, ps=None, actions=None, out_type='p'):
        own = np.array([np.asarray(o, dtype=np.float64).ravel()[:5] for o in obs])
        if self.single:
            out = []
            for i in range(self.N):
                if out_type.startswith('p'):
                    self.k[i] += 1
                    pi = script_pi(own[i:i + 1], self.k[i], done, self.tr.w[i:i + 1])[0]
                    self.tr.rec(2, i, float(bool(done)), np.asarray(obs[i], dtype=np.float32), pi)
                    out.append(pi)
                else:
                    v = script_v(own[i:i + 1], self.k[i])[0]
                    self.tr.rec(3, i, float(bool(done)), np.asarray(actions)[0, self.nbr[i]], v)
                    out.append(v)
            return np.array(out)[None]
        if out_type.startswith('p'):
            self.k += 1
            pi = script_pi(own, self.k, done, self.tr.w)
            self.tr.rec(2, float(bool(done)), np.array([np.asarray(o, dtype=np.float32).ravel() for o in obs]),
                        np.asarray(ps, dtype=np.float32)[0], pi)
            return pi[None]
        v = script_v(own, self.k)
        self.tr.rec(3, float(bool(done)), np.asarray(actions)[0], v)
        return v[None]

    def backward(self, obs_t, ps_t, acts_t, dones_t, Rs_t, Advs_t, lr, **kw):
        T = len(obs_t)
        acts, dones, Rs, Advs = acts_t[:, 0], dones_t[:, 0], Rs_t[:, 0], Advs_t[:, 0]
        if self.single:
            for i in range(self.N):
                o = np.array([np.asarray(obs_t[t][i], dtype=np.float32).ravel() for t in range(T)])
                self.tr.rec(5, i, lr, o, acts[:, self.nbr[i]], acts[:, i], dones, Rs[:, i], Advs[:, i])
        else:
            o = np.array([[np.asarray(obs_t[t][i], dtype=np.float32).ravel() for i in range(self.N)] for t in range(T)])
            self.tr.rec(5, lr, o, np.asarray(ps_t, dtype=np.float32)[:, 0], acts, dones, Rs, Advs)
        return {}


@pytest.mark.parametrize('name', CASES)
def test_oracle_agent_reproduces_reference_agent_class(name):
    g = golden(name)
    cp = load_cfg(str(g['ini']))
    env = OracleCACC(cp['ENV_CONFIG'])
    agent = OracleAgent(env.agent, env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma,
                        10 ** 6, cp['MODEL_CONFIG'], seed=12)
    tr = PolicyTrace(env.n_agent, env.n_a)
    nbr = [np.where(np.asarray(env.neighbor_mask)[i] == 1)[0] for i in range(env.n_agent)]
    agent.policy = ScriptedOraclePolicy(env.agent, tr, nbr)
    counter = Counter(int(g['total_step']), 10 ** 9, 10 ** 9)
    trainer = OracleTrainer(env, agent, counter)
    trainer.run()
    trace = np.array(tr.t)
    assert trace.shape == g['trace'].shape
    bad = np.flatnonzero(trace != g['trace'])
    assert bad.size == 0, (bad[:5], trace[bad[:5]], g['trace'][bad[:5]])
    got = np.array([[d['step'], d['avg_reward'], d['std_reward']] for d in trainer.data])
    np.testing.assert_array_equal(got, g['data'])
    assert env.seed == int(g['seed_after']) and counter.cur_step == int(g['cur_step'])
