"""CPU: the command line (main.py train / evaluate) wired end to end -- argument names of the reference, config
handling, directory layout, Counter / Trainer / Evaluator plumbing, checkpoint save / load calls -- with the CUDA
environment and agents replaced by the CPU oracle env and a scripted agent (the kernels are covered by the GPU tests)."""
import configparser
import os

import pytest

import main
from helpers import ROOT, ScriptedAgent
from oracle.cacc import OracleCACC


class _Env(OracleCACC):
    n_env = 1
    records = []

    def init_data(self, is_record, record_stats, output_path):
        self.records.append(('init_data', is_record, output_path))

    def collect_tripinfo(self):
        pass

    def output_data(self):
        self.records.append(('output_data',))


class _Agent(ScriptedAgent):
    def __init__(self, n_s_ls, n_a_ls, neighbor_mask, distance_mask, coop_gamma, total_step, config, seed=0, n_env=1):
        super().__init__('ma2c_nc', len(neighbor_mask), n_a_ls[0], config.getint('batch_size'))

    def save(self, model_dir, global_step):
        open(model_dir + 'checkpoint-%d.pt' % global_step, 'w').write('x')

    def load(self, model_dir, checkpoint=None):
        return any(f.startswith('checkpoint-') for f in os.listdir(model_dir))


@pytest.fixture
def cli(tmp_path, monkeypatch):
    cp = configparser.ConfigParser()
    cp.read(os.path.join(ROOT, 'config', 'config_ma2c_nc_catchup.ini'))
    cp['TRAIN_CONFIG']['total_step'] = '100'
    ini = tmp_path / 'exp.ini'
    with open(ini, 'w') as f:
        cp.write(f)
    monkeypatch.setattr(main, 'CACCEnv', _Env)
    monkeypatch.setitem(main.AGENTS, 'ma2c_nc', _Agent)
    monkeypatch.setattr(main.U, 'make_summary_writer', lambda d: None)
    monkeypatch.setattr(main.U, 'init_log', lambda d: None)
    _Env.records = []
    return str(tmp_path / 'run'), str(ini)


def test_reference_command_line_is_accepted():
    a = main.parse_args(['--base-dir', 'b', 'train', '--config-dir', 'c.ini'])
    assert (a.base_dir, a.option, a.config_dir) == ('b', 'train', 'c.ini')
    a = main.parse_args(['evaluate', '--evaluation-seeds', '2000,2010', '--demo'])
    assert (a.option, a.evaluation_seeds, a.demo, a.base_dir) == ('evaluate', '2000,2010', True, './runs/ma2c_nc_catchup')
    assert main.parse_args(['evaluate']).evaluation_seeds.split(',')[:2] == ['2000', '2010']
    with pytest.raises(SystemExit):
        main.parse_args([])


def test_train_then_evaluate(cli):
    base, ini = cli
    main.train(main.parse_args(['--base-dir', base, 'train', '--config-dir', ini]))
    assert sorted(os.listdir(base)) == ['data', 'log', 'model']
    assert sorted(os.listdir(base + '/data')) == ['exp.ini', 'train_reward.csv']
    rows = open(base + '/data/train_reward.csv').read().strip().split('\n')
    assert rows[0].split(',')[1:] == ['agent', 'step', 'test_id', 'avg_reward', 'std_reward'] and len(rows) == 2
    ckpt = os.listdir(base + '/model')
    assert len(ckpt) == 1 and ckpt[0].startswith('checkpoint-')           # saved at the final global step
    assert int(ckpt[0][len('checkpoint-'):-3]) == int(rows[1].split(',')[2])
    main.evaluate(main.parse_args(['--base-dir', base, 'evaluate', '--evaluation-seeds', '2000,2010']))
    assert ('init_data', True, base + '/eva_data/') in _Env.records and _Env.records[-1] == ('output_data',)


def test_evaluate_without_a_trained_agent_reports_and_returns(tmp_path, caplog, monkeypatch):
    monkeypatch.setattr(main.U, 'init_log', lambda d: None)
    main.evaluate(main.parse_args(['--base-dir', str(tmp_path / 'none'), 'evaluate', '--evaluation-seeds', '2000']))
    assert 'Cannot find .ini file' in caplog.text
