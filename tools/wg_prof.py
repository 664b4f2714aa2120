"""clock64 timeline of one CTA of the weight-gradient kernel (debug aid): MMA issuer per k-block and the accumulator
drains.  Usage: python tools/wg_prof.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from helpers import load_cfg
import main as M
from deeprl_network_b200 import _lib as L
from deeprl_network_b200.envs.cacc_env import CACCEnv
from deeprl_network_b200.utils import VecTrainer
cp = load_cfg('config_ma2c_nc_catchup.ini', n_env=4096)
env = CACCEnv(cp['ENV_CONFIG']); np.random.seed(12)
model = M.init_agent(env, cp['MODEL_CONFIG'], 10 ** 9, 12)
vt = VecTrainer(env, model, graph=False); vt.start()
for _ in range(2):
    vt.update()
e = model.engine
lib = L.lib(); lib.nmarl_debug_set_prof.argtypes = [C.c_void_p]
prof = torch.zeros(256, dtype=torch.int64, device='cuda')
e.rollout(env, sample='philox'); e.compute_returns()
lib.nmarl_debug_set_prof(prof.data_ptr())
e.backward(); torch.cuda.synchronize()
lib.nmarl_debug_set_prof(None)
p = prof.cpu().numpy(); t0 = p[0]
print('issuer (split 1, gate job 0, agent 1): per k-block [loop top, passes 1-2 issued, lo ready]')
for q in range(40):
    a, b, c = p[3 * q: 3 * q + 3]
    if a == 0: break
    print('  q=%2d top %8d  hi passes issued +%5d  lo ready +%5d   (since prev top %5d)' % (q, a - t0, b - a, c - b, a - p[3 * q - 3] if q else 0))
for name, base in (('A thread 0', 192), ('lo thread 0', 128)):
    print(name, 'drains [enter, accumulator ready, done]:')
    for s in range(5):
        a, b, c = p[base + 3 * s: base + 3 * s + 3]
        if a == 0: break
        print('  seg %d enter %8d  acc ready +%6d  rmw done +%5d' % (s, a - t0, b - a, c - b))
