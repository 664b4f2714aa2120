"""CPU oracle: agent wrappers + the rollout/train loop, B=1, with the reference's quirks.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PARITY UNPINNED for the parts that
sit on TensorFlow (see oracle/nets.py).  The rollout CONTROL FLOW (``OracleTrainer``, ``Counter``) is PINNED:
tests/test_trainer_flow.py replays traces recorded from the unmodified reference ``utils.Trainer`` driving a
scripted agent (tests/golden/trainer_*.npz) bit for bit.  It follows

  * ``IA2C`` / ``MA2C_*`` wrappers  agents/models.py:26-51, 198-227, 134-158, 246-258
  * ``Trainer._get_policy/_get_value/explore/perform/run``  utils.py:129-254
  * ``Counter``  utils.py:70-97

Quirks reproduced: Q1 value call re-runs the cell from the post-p state;
Q2 bootstrap advances LSTM state and consumes RNG; Q3 env seed steps by 2 per training
episode; Q4 logged reward comes from the interleaved greedy test episode; Q5 the global
counter counts training steps only; Q6 first batch has dones_bw[0]=False.

This is also the "restated reference (TF unavailable)" CPU baseline timed by bench.py.
"""
import itertools

import numpy as np
import torch

from .buffers import RolloutBuffer, Scheduler
from .nets import OraclePolicy


class OracleAgent:
    """Mirrors the method surface of agents/models.py (forward/add_transition/backward/reset)."""

    def __init__(self, variant, n_s_ls, n_a_ls, neighbor_mask, distance_mask, coop_gamma,
                 total_step, model_config, seed=0, params=None, dtype=torch.float32):
        g = lambda k: float(model_config[k])
        self.name = variant
        self.n_agent = len(neighbor_mask)
        self.n_step = int(model_config['batch_size'])
        self.reward_norm, self.reward_clip = g('reward_norm'), g('reward_clip')
        if variant == 'ia2c_fp':      # agents/models.py:172-177
            n_s_ls = [n + n_a_ls[0] * int(np.sum(np.asarray(neighbor_mask)[i])) for i, n in enumerate(n_s_ls)]
        self.policy = OraclePolicy(variant, n_s_ls, n_a_ls[0], neighbor_mask,
                                   n_h=int(model_config['num_lstm']), n_fc=int(model_config['num_fc']),
                                   params=params, dtype=dtype)
        self.neighbor_mask = np.asarray(neighbor_mask)
        if total_step:
            if model_config['lr_decay'] == 'constant':
                self.lr_scheduler = Scheduler(g('lr_init'), decay='constant')
            else:
                self.lr_scheduler = Scheduler(g('lr_init'), g('lr_min'), total_step, decay=model_config['lr_decay'])
            self.hp = dict(v_coef=g('value_coef'), e_coef=g('entropy_coef'), max_grad_norm=g('max_grad_norm'),
                           alpha=g('rmsp_alpha'), epsilon=g('rmsp_epsilon'))
            self.trans_buffer = RolloutBuffer(g('gamma'), coop_gamma, distance_mask)
        self.last_summary = None

    def forward(self, obs, done, ps_or_nactions=None, actions=None, out_type='p'):
        if self.name.startswith('ia2c'):
            # IA2C signature: forward(obs, done, nactions=None, out_type='p') (models.py:44-51)
            if isinstance(actions, str):
                out_type, actions = actions, None
            if out_type.startswith('p'):
                out = self.policy.forward(obs, done, None, None, 'p')[0]
            else:
                a = np.zeros((1, self.n_agent), dtype=np.int64)
                for i in range(self.n_agent):     # scatter neighbour actions back to a full vector
                    for k, j in enumerate(np.where(self.neighbor_mask[i] == 1)[0]):
                        a[0, j] = ps_or_nactions[i][k]
                out = self.policy.forward(obs, done, None, a, 'v')[0]
            return [out[i] for i in range(self.n_agent)]
        ps = np.asarray(ps_or_nactions)[None]
        a = None if actions is None else np.asarray(actions)[None]
        return self.policy.forward(obs, done, ps, a, out_type)[0]

    def add_transition(self, ob, p, action, reward, value, done):
        if self.reward_norm > 0:
            reward = reward / self.reward_norm
        if self.reward_clip > 0:
            reward = np.clip(reward, -self.reward_clip, self.reward_clip)
        if self.name.startswith('ia2c'):
            p = np.zeros((self.n_agent, self.policy.n_a))    # unused placeholder
        self.trans_buffer.add_transition([np.asarray(o) for o in ob], np.array(p), np.asarray(action),
                                         reward, np.asarray(value), done)

    def backward(self, Rends, dt=0, summary_writer=None, global_step=None, apply=True):
        lr = self.lr_scheduler.get(self.n_step)
        buf = self.trans_buffer
        obs_t = [[np.asarray(o)[None] for o in ob] for ob in buf.obs]         # T x N x [1,n_s_i]
        ps_t = np.array(buf.adds, dtype=np.float32)[:, None]                   # [T,1,N,n_a]
        acts_t = np.array(buf.acts, dtype=np.int64)[:, None]                    # [T,1,N]
        dones, Rs, Advs = buf.finish(Rends)
        Rs_t = np.transpose(Rs)[:, None]                                         # [T,1,N]
        Advs_t = np.transpose(Advs)[:, None]
        dones_t = dones.astype(np.float64)[:, None]
        self.last_batch = dict(Rs=Rs, Advs=Advs, dones=dones)
        self.last_summary = self.policy.backward(obs_t, None if self.name.startswith('ia2c') else ps_t, acts_t, dones_t,
                                                 Rs_t, Advs_t, lr, apply=apply, **self.hp)
        self.last_summary['lr'] = lr
        return self.last_summary

    def reset(self):
        self.policy.reset()


class Counter:
    """utils.py:70-97"""

    def __init__(self, total_step, test_step, log_step):
        self.counter = itertools.count(1)
        self.cur_step = 0
        self.cur_test_step = 0
        self.total_step, self.test_step, self.log_step = total_step, test_step, log_step
        self.stop = False

    def next(self):
        self.cur_step = next(self.counter)
        return self.cur_step

    def should_stop(self):
        return self.cur_step >= self.total_step or self.stop


class OracleTrainer:
    """utils.py:100-254 without TF summaries / CSV.  ``uniform_fn`` (optional) supplies the
    uniform used for each action draw so tests can feed identical randomness to the CUDA path;
    by default it is ``np.random.random_sample`` == what ``np.random.choice`` consumes."""

    def __init__(self, env, model, counter, uniform_fn=None):
        self.env, self.model, self.global_counter = env, model, counter
        self.agent = env.agent
        self.n_step = model.n_step
        assert env.T % self.n_step == 0          # utils.py:110
        self.env.train_mode = True
        self.uniform_fn = uniform_fn or np.random.random_sample
        self.data = []
        self.trace = []

    @staticmethod
    def choice(p, u):
        """np.random.choice(n, p=p) for one draw: cdf.searchsorted(u, 'right')."""
        cdf = np.cumsum(np.asarray(p, dtype=np.float64))
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, u, side='right'))

    def _get_policy(self, ob, done, mode='train'):
        if self.agent.startswith('ma2c'):
            self.ps = self.env.get_fingerprint()
            policy = self.model.forward(ob, done, self.ps)
        else:
            policy = self.model.forward(ob, done)
        action = []
        for pi in policy:
            if mode == 'train':
                action.append(self.choice(pi, self.uniform_fn()))
            else:
                action.append(int(np.argmax(pi)))
        return policy, np.array(action)

    def _get_value(self, ob, done, action):
        if self.agent.startswith('ma2c'):
            return self.model.forward(ob, done, self.ps, np.array(action), 'v')
        self.naction = self.env.get_neighbor_action(action)
        return self.model.forward(ob, done, self.naction, 'v')

    def explore(self, prev_ob, prev_done):
        ob, done = prev_ob, prev_done
        for _ in range(self.n_step):
            policy, action = self._get_policy(ob, done)
            value = self._get_value(ob, done, action)
            self.env.update_fingerprint(policy)
            next_ob, reward, done, global_reward = self.env.step(action)
            self.episode_rewards.append(global_reward)
            self.global_counter.next()
            self.cur_step += 1
            self.model.add_transition(ob, self.ps if self.agent.startswith('ma2c') else self.naction,
                                      action, reward, value, done)
            if done:
                break
            ob = next_ob
        if done:
            R = np.zeros(self.model.n_agent)
        else:
            _, action = self._get_policy(ob, done)
            R = self._get_value(ob, done, action)
        return ob, done, np.asarray(R)

    def perform(self, test_ind):
        ob = self.env.reset(test_ind=test_ind)
        rewards = []
        done = True
        self.model.reset()
        while True:
            policy, action = self._get_policy(ob, done, mode='test')
            self.env.update_fingerprint(policy)
            next_ob, reward, done, global_reward = self.env.step(action)
            rewards.append(global_reward)
            if done:
                break
            ob = next_ob
        return np.mean(np.array(rewards)), np.std(np.array(rewards))

    def run(self, max_episodes=None, test_episode=True):
        n_ep = 0
        while not self.global_counter.should_stop():
            ob = self.env.reset()
            done = True
            self.model.reset()
            self.cur_step = 0
            self.episode_rewards = []
            while True:
                ob, done, R = self.explore(ob, done)
                dt = self.env.T - self.cur_step
                summ = self.model.backward(R, dt)
                self.trace.append(summ)
                if done:
                    self.env.terminate()
                    break
            mean_reward, std_reward = np.mean(self.episode_rewards), np.std(self.episode_rewards)
            if test_episode:
                self.env.train_mode = False
                mean_reward, std_reward = self.perform(-1)
                self.env.train_mode = True
            self.data.append(dict(agent=self.agent, step=self.global_counter.cur_step, test_id=-1,
                                  avg_reward=mean_reward, std_reward=std_reward))
            n_ep += 1
            if max_episodes is not None and n_ep >= max_episodes:
                break
        return self.data
