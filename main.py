"""Train / evaluate networked-MARL agents on the B200-native hot path.

Same command line as the reference (main.py:21-40):
    python main.py --base-dir D train --config-dir F
    python main.py --base-dir D evaluate [--evaluation-seeds s1,s2] [--demo]
and the same .ini surface (MODEL_CONFIG / TRAIN_CONFIG / ENV_CONFIG).  Optional new keys:
ENV_CONFIG.n_env (parallel episodes per process; > 1 selects the batched VecTrainer).
Only the CACC scenarios and the agents on the hot path (ia2c, ma2c_nc, ma2c_ic3, ma2c_dial)
are available; ATSC/SUMO environments are out of scope.
"""
import argparse
import configparser
import logging

from deeprl_network_b200.agents.models import IA2C, IA2C_FP, IA2C_CU, MA2C_NC, MA2C_IC3, MA2C_DIAL
from deeprl_network_b200.envs.cacc_env import CACCEnv
from deeprl_network_b200.utils import (Counter, Trainer, Evaluator, VecTrainer, check_dir, copy_file, find_file,
                                       init_dir, init_log, make_summary_writer)

AGENTS = {'ia2c': IA2C, 'ia2c_fp': IA2C_FP, 'ma2c_cu': IA2C_CU, 'ma2c_nc': MA2C_NC, 'ma2c_ic3': MA2C_IC3, 'ma2c_dial': MA2C_DIAL}


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument('--base-dir', type=str, required=False, default='./runs/ma2c_nc_catchup',
                        help="experiment base dir")
    subparsers = parser.add_subparsers(dest='option', help="train or evaluate")
    sp = subparsers.add_parser('train', help='train a single agent under base dir')
    sp.add_argument('--config-dir', type=str, required=False, default='./config/config_ma2c_nc_catchup.ini',
                    help="experiment config path")
    sp = subparsers.add_parser('evaluate', help="evaluate and compare agents under base dir")
    sp.add_argument('--evaluation-seeds', type=str, required=False,
                    default=','.join([str(i) for i in range(2000, 2500, 10)]),
                    help="random seeds for evaluation, split by ,")
    sp.add_argument('--demo', action='store_true', help="no-op here (SUMO gui in the reference)")
    args = parser.parse_args()
    if not args.option:
        parser.print_help()
        exit(1)
    return args


def init_env(config, port=0):
    scenario = config.get('scenario')
    if scenario.startswith('atsc'):
        raise NotImplementedError('ATSC/SUMO environments are outside the accelerated hot path')
    return CACCEnv(config)


def init_agent(env, config, total_step, seed, **kw):
    cls = AGENTS.get(env.agent)
    if cls is None:
        logging.error('agent %r is not on the accelerated hot path' % env.agent)
        return None
    if env.agent == 'ia2c' and env.n_env > 1:     # device-resident rollouts gather neighbour observations in the kernel
        kw.setdefault('obs_mode', 'gather')
    return cls(env.n_s_ls, env.n_a_ls, env.neighbor_mask, env.distance_mask, env.coop_gamma,
               total_step, config, seed=seed, n_env=env.n_env, **kw)


def train(args):
    dirs = init_dir(args.base_dir)
    init_log(dirs['log'])
    copy_file(args.config_dir, dirs['data'])
    config = configparser.ConfigParser()
    config.read(args.config_dir)
    env = init_env(config['ENV_CONFIG'])
    logging.info('Training: a dim %r, agent dim: %d' % (env.n_a_ls, env.n_agent))
    total_step = int(config.getfloat('TRAIN_CONFIG', 'total_step'))
    test_step = int(config.getfloat('TRAIN_CONFIG', 'test_interval'))
    log_step = int(config.getfloat('TRAIN_CONFIG', 'log_interval'))
    global_counter = Counter(total_step, test_step, log_step)
    seed = config.getint('ENV_CONFIG', 'seed')
    model = init_agent(env, config['MODEL_CONFIG'], total_step, seed)
    summary_writer = make_summary_writer(dirs['log'])
    if env.n_env == 1:
        Trainer(env, model, global_counter, summary_writer, output_path=dirs['data']).run()
        final_step = global_counter.cur_step
    else:
        vt = VecTrainer(env, model)
        vt.start()
        steps = 0
        while steps < total_step:
            vt.update()
            steps += model.n_step * env.n_env
            if vt.n_update % 10 == 0:
                logging.info('update %d, env steps %d, mean step reward %.2f' % (vt.n_update, steps, vt.mean_reward()))
        final_step = steps
    logging.info('Training: save final model at step %d ...' % final_step)
    model.save(dirs['model'], final_step)


def evaluate_fn(agent_dir, output_dir, seeds, port, demo):
    agent = agent_dir.split('/')[-1]
    if not check_dir(agent_dir):
        logging.error('Evaluation: %s does not exist!' % agent)
        return
    config_dir = find_file(agent_dir + '/data/')
    if not config_dir:
        return
    config = configparser.ConfigParser()
    config.read(config_dir)
    config['ENV_CONFIG']['n_env'] = '1'
    env = init_env(config['ENV_CONFIG'], port=port)
    env.init_test_seeds(seeds)
    model = init_agent(env, config['MODEL_CONFIG'], 0, 0)
    if model is None:
        return
    if not model.load(agent_dir + '/model/'):
        return
    Evaluator(env, model, output_dir, gui=demo).run()


def evaluate(args):
    base_dir = args.base_dir
    if not args.demo:
        dirs = init_dir(base_dir, pathes=['eva_data', 'eva_log'])
        init_log(dirs['eva_log'])
        output_dir = dirs['eva_data']
    else:
        output_dir = None
    seeds = args.evaluation_seeds
    logging.info('Evaluation: random seeds: %s' % seeds)
    seeds = [int(s) for s in seeds.split(',')] if seeds else []
    evaluate_fn(base_dir, output_dir, seeds, 1, args.demo)


if __name__ == '__main__':
    args = parse_args()
    if args.option == 'train':
        train(args)
    else:
        evaluate(args)
