import sys, os, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, ctypes as C
from gpu_common import make_pair
from test_gpu_backward import _batch
from deeprl_network_b200 import _lib as L
T, B = 3, 128
eng, orc, lay, params = make_pair('ma2c_nc', B, T=T, dtype=torch.float64)
batch = _batch(eng, lay, T, B)
a = eng._bwd_args(T)
eng.h_seq[0].copy_(eng.h_bw); eng.c_seq[0].copy_(eng.c_bw)
print('train fwd...', flush=True)
L.check(L.lib().nmarl_a2c_train_forward(C.byref(eng.model), C.byref(a), L.stream()), 'fwd')
torch.cuda.synchronize(); print('ok, err', eng.tc_err.item(), flush=True)
print('bptt...', flush=True)
L.check(L.lib().nmarl_a2c_bptt(C.byref(eng.model), C.byref(a), L.stream()), 'bptt')
torch.cuda.synchronize(); print('ok, err', eng.tc_err.item(), flush=True)
print(float(eng.grads.abs().sum()))
