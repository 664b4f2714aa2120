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
