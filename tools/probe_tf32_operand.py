"""Hardware probes (run on a B200; result of round 2, session r2a: profiles/r2a_probe.txt -- both answers were "bit-identical"
when the packed operands were still split by truncation.  The product now splits by round-to-nearest (tc.cuh), so (1)
reports a small difference today: the raw tile is read truncated by the tensor core, the packed hi tile is rounded):
    python tools/probe_tf32_operand.py
(1) Does the TF32 tensor-core datapath ignore the 13 low mantissa bits of a shared-memory operand?  The stand-alone
    3xTF32 GEMM runs with hi tile = masked values (what the product packs) and with hi tile = raw fp32 values; the
    accumulators must be bit-identical for single-copy operand tiles to be exact.
(2) Does the both-operands-in-shared-memory form (A staged by the row threads into a 128B-swizzled tile, made visible
    with fence.proxy.async) reproduce the A-in-TMEM result?  That is the form a two-tiles-per-SM kernel needs."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprl_network_b200 import _lib as L  # noqa: E402


def main():
    lib = L.lib()
    args = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 3
    lib.nmarl_tc_gemm_selftest.argtypes = args
    lib.nmarl_tc_gemm_selftest_raw.argtypes = args
    lib.nmarl_tc_gemm_selftest_ss.argtypes = args
    torch.manual_seed(0)
    for (M, K, N) in [(256, 256, 256), (128, 64, 64), (512, 192, 256)]:
        a = torch.randn(M, K, device='cuda') * torch.logspace(-3, 3, K, device='cuda')
        w = torch.randn(K, N, device='cuda')
        out = []
        for fn in (lib.nmarl_tc_gemm_selftest, lib.nmarl_tc_gemm_selftest_raw, lib.nmarl_tc_gemm_selftest_ss):
            c = torch.zeros(M, N, device='cuda')
            scratch = torch.zeros(((K + 31) // 32) * 2 * N * 32, device='cuda')
            err = torch.zeros(1, dtype=torch.int32, device='cuda')
            L.check(fn(a.data_ptr(), w.data_ptr(), c.data_ptr(), M, K, N, scratch.data_ptr(), err.data_ptr(), L.stream()), 'selftest')
            torch.cuda.synchronize()
            if int(err.item()) != 0:
                print('  %s: watchdog code %d' % (fn.__name__, int(err.item())))
            out.append(c)
        ref = (a.double() @ w.double())
        same = torch.equal(out[0], out[1])
        print('M=%d K=%d N=%d: raw-operand GEMM %s the masked one; max rel err masked %.2e raw %.2e' % (
            M, K, N, 'bit-identical to' if same else 'DIFFERS from',
            float(((out[0] - ref).abs().max() / ref.abs().max())), float(((out[1] - ref).abs().max() / ref.abs().max()))))
        print('    shared-memory A operand: %s the TMEM-A result (max rel err %.2e)' % (
            'bit-identical to' if torch.equal(out[0], out[2]) else 'DIFFERS from',
            float(((out[2] - ref).abs().max() / ref.abs().max()))))


if __name__ == '__main__':
    main()
