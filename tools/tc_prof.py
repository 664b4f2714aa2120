"""Per-phase clock64() timeline of one CTA of the tensor-core forward kernel (debug aid)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from gpu_common import make_pair
from deeprl_network_b200 import _lib as L
B = 4096
eng, orc, lay, _ = make_pair('ma2c_nc', B)
lib = L.lib()
prof = torch.zeros(256, dtype=torch.int64, device='cuda')
obs = torch.randn(8, B, 8, device='cuda'); fp = torch.softmax(torch.randn(8, B, 4, device='cuda'), -1); done = torch.zeros(B, device='cuda')
pi = torch.zeros(8, B, 4, device='cuda'); act = torch.zeros(8, B, dtype=torch.int32, device='cuda'); v = torch.zeros(8, B, device='cuda')
for it in range(3):
    eng.step_p(obs, fp, done, pi, act, L.SAMPLE_PHILOX)
lib.nmarl_debug_set_prof.argtypes = [C.c_void_p]
mode = sys.argv[1] if len(sys.argv) > 1 else 'p'
if mode == 'ps':                       # rollout p-call that also saves the BPTT activations
    eng._alloc_train()
    eng.h_seq[0].copy_(eng.h[eng.cur]); eng.c_seq[0].copy_(eng.c[eng.cur])
    for it in range(2):
        eng._seq_call(0, obs, fp, done, 'p', pi=pi, action=act, mode=L.SAMPLE_PHILOX, save=True)
    lib.nmarl_debug_set_prof(prof.data_ptr())
    eng._seq_call(0, obs, fp, done, 'p', pi=pi, action=act, mode=L.SAMPLE_PHILOX, save=True)
else:
    lib.nmarl_debug_set_prof(prof.data_ptr())
    eng.step_p(obs, fp, done, pi, act, L.SAMPLE_PHILOX)
torch.cuda.synchronize()
print('mode', mode)
p = prof.cpu().numpy()
n = int(p[31])
names = ['start', 'inputs loaded', 'enc A produced', 'enc ready', 'all A produced', 'acc ready', 'it0 tmem loaded', 'it0 math done',
         'it0 stored', 'it0 heads', 'it1 tmem loaded', 'it1 math done', 'it1 stored', 'it1 heads', 'cell done', 'heads barrier', 'end']
t0 = p[0]
print('row thread (agent 1, tile 0):')
for i in range(n):
    print('  %-28s %8d cyc (+%d)' % (names[i] if i < len(names) else '?', p[i] - t0, p[i] - p[i - 1] if i else 0))
print('MMA thread: per k-block [b ready, a ready, issued]')
for q in range(14):
    a, b, c = p[32 + 3 * q: 35 + 3 * q]
    if a == 0: break
    print('  q=%2d  wait_b done %8d  wait_a done %8d (+%d)  issued %8d (+%d)' % (q, a - t0, b - t0, b - a, c - t0, c - b))
lib.nmarl_debug_set_prof(None)
