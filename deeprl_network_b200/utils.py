"""Rollout / train loop and experiment plumbing -- mirror of the reference's utils.py.

``Counter``, ``Trainer`` (``explore`` / ``perform`` / ``run``), ``Evaluator`` and the directory /
logging helpers keep the reference's names, arguments and control flow (utils.py:11-60, 70-97,
100-254, 311-336), including its quirks: the value call after the policy
call, the state-advancing bootstrap, the interleaved greedy test episode for CACC whose
reward is what gets logged, and the counter that only counts training steps.

``VecTrainer`` is the new batched loop (n_env parallel episodes, everything device resident,
optionally one process per GPU with one NCCL gradient all-reduce per update).
"""
import itertools
import logging
import os
import shutil
import time

import numpy as np
import torch


def check_dir(cur_dir):
    return os.path.exists(cur_dir)


def copy_file(src_dir, tar_dir):
    shutil.copy(src_dir, tar_dir)


def find_file(cur_dir, suffix='.ini'):
    for file in os.listdir(cur_dir):
        if file.endswith(suffix):
            return cur_dir + '/' + file
    logging.error('Cannot find %s file' % suffix)
    return None


def init_dir(base_dir, pathes=['log', 'data', 'model']):
    if not os.path.exists(base_dir):
        os.mkdir(base_dir)
    dirs = {}
    for path in pathes:
        cur_dir = base_dir + '/%s/' % path
        if not os.path.exists(cur_dir):
            os.mkdir(cur_dir)
        dirs[path] = cur_dir
    return dirs


def init_log(log_dir):
    logging.basicConfig(format='%(asctime)s [%(levelname)s] %(message)s', level=logging.INFO,
                        handlers=[logging.FileHandler('%s/%d.log' % (log_dir, time.time())), logging.StreamHandler()])


def init_test_flag(test_mode):
    return {'no_test': (False, False), 'in_train_test': (True, False), 'after_train_test': (False, True),
            'all_test': (True, True)}.get(test_mode, (False, False))


def make_summary_writer(log_dir):
    """TensorBoard event writer (replaces tf.summary.FileWriter); None if tensorboard is absent."""
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir)
    except Exception:  # pragma: no cover
        logging.warning('tensorboard not available: scalar summaries disabled')
        return None


class Counter:
    """utils.py:70-97"""

    def __init__(self, total_step, test_step, log_step):
        self.counter = itertools.count(1)
        self.cur_step = 0
        self.cur_test_step = 0
        self.total_step = total_step
        self.test_step = test_step
        self.log_step = log_step
        self.stop = False

    def next(self):
        self.cur_step = next(self.counter)
        return self.cur_step

    def should_test(self):
        test = False
        if (self.cur_step - self.cur_test_step) >= self.test_step:
            test = True
            self.cur_test_step = self.cur_step
        return test

    def should_log(self):
        return self.cur_step % self.log_step == 0

    def should_stop(self):
        if self.cur_step >= self.total_step:
            return True
        return self.stop


class Trainer:
    """One environment, one episode at a time, exactly like the reference (utils.py:100-254)."""

    def __init__(self, env, model, global_counter, summary_writer, output_path=None, uniform_fn=None):
        self.cur_step = 0
        self.global_counter = global_counter
        self.env = env
        self.agent = self.env.agent
        self.model = model
        self.sess = getattr(model, 'sess', None)
        self.n_step = self.model.n_step
        self.summary_writer = summary_writer
        assert self.env.T % self.n_step == 0
        self.data = []
        self.output_path = output_path
        self.env.train_mode = True
        # the uniform each np.random.choice draw consumes (tests inject their own stream)
        self.uniform_fn = uniform_fn

    def _add_summary(self, reward, global_step, is_train=True):
        if self.summary_writer is not None:
            self.summary_writer.add_scalar('train_reward' if is_train else 'test_reward', reward, global_step)

    def _sample(self, pi):
        if self.uniform_fn is None:
            return np.random.choice(np.arange(len(pi)), p=pi)
        cdf = np.cumsum(np.asarray(pi, dtype=np.float64))
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, self.uniform_fn(), side='right'))

    def _get_policy(self, ob, done, mode='train'):
        if self.agent.startswith('ma2c'):
            self.ps = self.env.get_fingerprint()
            policy = self.model.forward(ob, done, self.ps)
        else:
            policy = self.model.forward(ob, done)
        action = []
        for pi in policy:
            action.append(self._sample(pi) if mode == 'train' else np.argmax(pi))
        return policy, np.array(action)

    def _get_value(self, ob, done, action):
        if self.agent.startswith('ma2c'):
            return self.model.forward(ob, done, self.ps, np.array(action), 'v')
        self.naction = self.env.get_neighbor_action(action)
        if not self.naction:
            self.naction = np.nan
        return self.model.forward(ob, done, self.naction, 'v')

    def _log_episode(self, global_step, mean_reward, std_reward):
        self.data.append({'agent': self.agent, 'step': global_step, 'test_id': -1,
                          'avg_reward': mean_reward, 'std_reward': std_reward})
        self._add_summary(mean_reward, global_step)
        if self.summary_writer is not None:
            self.summary_writer.flush()

    def explore(self, prev_ob, prev_done):
        ob, done = prev_ob, prev_done
        for _ in range(self.n_step):
            policy, action = self._get_policy(ob, done)          # pre-decision
            value = self._get_value(ob, done, action)            # post-decision (quirk Q1)
            self.env.update_fingerprint(policy)
            next_ob, reward, done, global_reward = self.env.step(action)
            self.episode_rewards.append(global_reward)
            global_step = self.global_counter.next()
            self.cur_step += 1
            if self.agent.startswith('ma2c'):
                self.model.add_transition(ob, self.ps, action, reward, value, done)
            else:
                self.model.add_transition(ob, self.naction, action, reward, value, done)
            if self.global_counter.should_log():
                logging.info('''Training: global step %d, episode step %d,
                                   ob: %s, a: %s, pi: %s, r: %.2f, train r: %.2f, done: %r''' %
                             (global_step, self.cur_step, str(ob), str(action), str(policy), global_reward,
                              np.mean(reward), done))
            if done:                                             # terminal check inside the batch loop
                break
            ob = next_ob
        if done:
            R = np.zeros(self.model.n_agent)
        else:                                                    # quirk Q2
            _, action = self._get_policy(ob, done)
            R = self._get_value(ob, done, action)
        return ob, done, R

    def perform(self, test_ind, gui=False):
        ob = self.env.reset(gui=gui, test_ind=test_ind)
        rewards = []
        done = True                                              # pre-decision done resets the LSTM
        self.model.reset()
        while True:
            if self.env.name.startswith('atsc'):
                policy, action = self._get_policy(ob, done)
            else:                                                # CACC: deterministic test policy
                policy, action = self._get_policy(ob, done, mode='test')
            self.env.update_fingerprint(policy)
            next_ob, reward, done, global_reward = self.env.step(action)
            rewards.append(global_reward)
            if done:
                break
            ob = next_ob
        return np.mean(np.array(rewards)), np.std(np.array(rewards))

    def run(self, max_episodes=None):
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
                global_step = self.global_counter.cur_step
                self.model.backward(R, dt, self.summary_writer, global_step)
                if done:
                    self.env.terminate()
                    break
            rewards = np.array(self.episode_rewards)
            mean_reward, std_reward = np.mean(rewards), np.std(rewards)
            if not self.env.name.startswith('atsc'):             # quirk Q4
                self.env.train_mode = False
                mean_reward, std_reward = self.perform(-1)
                self.env.train_mode = True
            self._log_episode(global_step, mean_reward, std_reward)
            n_ep += 1
            if max_episodes is not None and n_ep >= max_episodes:
                break
        if self.output_path is not None:
            import pandas as pd
            pd.DataFrame(self.data).to_csv(self.output_path + 'train_reward.csv')


class Tester(Trainer):
    """Imported by the reference's main.py but never invoked; kept for API parity."""

    def __init__(self, env, model, global_counter, summary_writer, output_path):
        super().__init__(env, model, global_counter, summary_writer)
        self.env.train_mode = False
        self.test_num = self.env.test_num
        self.output_path = output_path
        self.data = []


class Evaluator(Tester):
    """utils.py:311-336"""

    def __init__(self, env, model, output_path, gui=False):
        self.env = env
        self.model = model
        self.agent = self.env.agent
        self.env.train_mode = False
        self.test_num = self.env.test_num
        self.output_path = output_path
        self.gui = gui
        self.uniform_fn = None

    def run(self):
        is_record = not self.gui
        self.env.cur_episode = 0
        self.env.init_data(is_record, False, self.output_path)
        for test_ind in range(self.test_num):
            reward, _ = self.perform(test_ind, gui=self.gui)
            self.env.terminate()
            logging.info('test %i, avg reward %.2f' % (test_ind, reward))
            self.env.collect_tripinfo()
        self.env.output_data()


class VecTrainer:
    """Batched training loop: n_env parallel episodes advance in lock-step on the device.

    Per update: ``rollout`` (n_step x [p-call, v-call, env step] + bootstrap) -> returns ->
    training forward/BPTT/wgrad -> [all-reduce] -> clip + RMSProp, then per-env auto-reset of the
    environments whose episode ended (model.reset() + env.reset() of the reference, per env).
    With ``graph=True`` one update is captured once into a CUDA graph and replayed.
    """

    def __init__(self, env, model, graph=True, sample='philox'):
        self.env, self.model, self.engine = env, model, model.engine
        assert env.n_env == model.n_env
        self.sample = sample
        self.use_graph = graph
        self.graph = None
        self.n_update = 0
        self.env.train_mode = True

    def start(self):
        self._seed = self.env.seed
        self.env.reset_device(u01=None, philox_seed=self._seed)
        self.engine.reset_states()
        self.engine.begin_episode(self.env)

    def _one_update(self, uniforms=None):
        e, env = self.engine, self.env
        e.rollout(env, sample=self.sample, uniforms=uniforms)
        e.update(self._lr)
        # episode boundaries: envs whose last step returned done restart (per-env model.reset/env.reset)
        done = e.done_buf[e.T_cur]
        e.roll_buffers()
        e.reset_states(mask=done)
        env.reset_device(u01=None, mask=done, obs_out=e.obs_buf[0], fp_out=e.fp_buf[0], philox_seed=self._seed)
        e.normalize_cur()

    def update(self, uniforms=None):
        e = self.engine
        lr = self.model.lr_scheduler.get(self.model.n_step)
        e.lr_dev.fill_(float(lr))
        self._lr = e.lr_dev
        if not self.use_graph:
            self._one_update(uniforms)
        else:
            if self.graph is None:
                self._static_uniforms = uniforms
                # warm-up outside capture (sets kernel attributes, allocates training buffers)
                self._one_update(uniforms)
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._one_update(self._static_uniforms)
                self.n_update += 1
                return
            self.graph.replay()
        self.n_update += 1

    def mean_reward(self):
        """Mean per-step global reward of the last batch (host sync)."""
        return float(self.engine.grew_buf[:self.engine.T_cur].mean().item())
