// tc_wgrad.cu -- tcgen05 weight gradient of the LSTM gate matrix:
//   dW[ka][n] = sum over all (t, env) rows r of  [s | h^][r][ka] * dz[r][n]        (ka, n in [0, 256))
// A 3xTF32 GEMM with M = ka (two 128-lane tiles), N = 256 and a contraction over rows split across CTAs
// (fixed-order reduce afterwards).  The A operand is read straight from the row-major saved activations:
// TMEM lane = ka, so for a fixed row the 32 lanes of a warp read 32 consecutive floats (coalesced), split
// them hi/lo and tcgen05.st them.  The B operand (dz^T, K-major over rows) was written by the backward cell
// kernel as ready-made [hi | lo] 128B-swizzled tiles, so the producer is one 64 KB bulk copy per 32 rows.
#include "bwd_common.cuh"
#include "tc_row.cuh"

namespace {
using namespace tcrow;

struct TcWgK {
  int N, B, T, splits, i_lda;
  const float* A;            // sv_sh   [T][N][B][lda]
  const float* BT;           // dzT     [T][N][B/32][2][256][32]
  float* ws;                 // [splits][N][257][256]
  int* err;
};

__global__ void __launch_bounds__(TC_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ TcWgK k) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bst = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_STAGES * STAGE_BYTES);
  uint64_t* b_full = bars, *b_empty = bars + S_STAGES, *a_full = bars + 2 * S_STAGES, *a_empty = a_full + 2;
  uint64_t* enc_full = a_empty + 2, *acc_full = enc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int mt = blockIdx.x & 1, sp = blockIdx.x >> 1, i = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kb_total = k.T * (k.B / 32);                              // 32-row k-blocks of this agent
  const int per = (kb_total + k.splits - 1) / k.splits;
  const int kb0 = sp * per, kb1 = min(kb_total, kb0 + per);
  const int nkb = max(0, kb1 - kb0);
  const int bpt = k.B / 32;                                           // k-blocks per time step

  if (tid == 0) {
    for (int s = 0; s < S_STAGES; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&a_full[s], ROW_THREADS); tc::mbar_init(&a_empty[s], 1); }
    tc::mbar_init(enc_full, 1);
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == ROW_THREADS / 32 + 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < ROW_THREADS / 32) {
    RowCtx c;
    const int set = warp >> 2, quarter = warp & 3;
    const int ka = mt * 128 + quarter * 32 + lane;
    c.tmem = tmem; c.lane_base = (uint32_t)(quarter * 32) << 16;
    c.a_full = a_full; c.a_empty = a_empty; c.enc_full = enc_full; c.q = 0; c.e = 0; c.set = set; c.err = k.err;
    for (int kb = kb0; kb < kb1; ++kb) {
      const int t = kb / bpt, rb = kb - t * bpt;
      const float* src = k.A + (((size_t)t * k.N + i) * k.B + rb * 32 + set * W) * k.i_lda + ka;
      float x[W];
#pragma unroll
      for (int j = 0; j < W; ++j) x[j] = src[(size_t)j * k.i_lda];
      produce_in(c, x);
    }
    float* out = k.ws + (((size_t)sp * k.N + i) * 257 + ka) * 256 + set * 64;
    if (nkb > 0) {
      tc::mbar_wait(acc_full, 0, k.err, 13);
      tc::fence_after_sync();
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        float v[8];
        tc::tmem_ld8(tmem + c.lane_base + ACC_COL + set * 64 + 8 * p, v);
        tc::wait_ld();
        store_vec<8>(out + 8 * p, v);
      }
      tc::fence_before_sync();
    } else {
      float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int p = 0; p < 8; ++p) store_vec<8>(out + 8 * p, z);
    }
  } else if (warp == ROW_THREADS / 32) {
    if (lane == 0) {
      for (int q = 0; q < nkb; ++q) {
        const int kb = kb0 + q, st = q % S_STAGES;
        const int t = kb / bpt, rb = kb - t * bpt;
        tc::mbar_wait(&b_empty[st], ((q / S_STAGES) & 1) ^ 1, k.err, 21);
        tc::mbar_arrive_expect_tx(&b_full[st], STAGE_BYTES);
        tc::bulk_g2s(bst + st * STAGE_BYTES, k.BT + (((size_t)t * k.N + i) * bpt + rb) * (2 * 256 * 32), STAGE_BYTES, &b_full[st]);
      }
    }
  } else {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(128, 256);
      for (int q = 0; q < nkb; ++q) {
        const int st = q % S_STAGES, slot = q & 1;
        tc::mbar_wait(&b_full[st], (q / S_STAGES) & 1, k.err, 31);
        tc::mbar_wait(&a_full[slot], (q >> 1) & 1, k.err, 32);
        tc::fence_after_sync();
        const uint64_t d_hi = tc::smem_desc_sw128(bst + st * STAGE_BYTES), d_lo = tc::smem_desc_sw128(bst + st * STAGE_BYTES + 256 * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a_hi = tmem + A_COL + slot * 64 + ks * 8, a_lo = a_hi + 32;
          tc::mma_tf32_ts(tmem + ACC_COL, a_hi, d_hi + 2 * ks, idesc, (q == 0 && ks == 0) ? 0u : 1u);
          tc::mma_tf32_ts(tmem + ACC_COL, a_hi, d_lo + 2 * ks, idesc, 1u);
          tc::mma_tf32_ts(tmem + ACC_COL, a_lo, d_hi + 2 * ks, idesc, 1u);
        }
        tc::mma_commit(&a_empty[slot]);
        tc::mma_commit(&b_empty[st]);
      }
      if (nkb > 0) tc::mma_commit(acc_full);
    }
  }
  __syncthreads();
  if (warp == ROW_THREADS / 32 + 1) { tc::fence_after_sync(); tc::tmem_dealloc(tmem, 512); }
}

// column sums of dz (the gate bias gradient): [splits][N][256] partials, fixed-order
__global__ void __launch_bounds__(256) dz_colsum_kernel(const float* __restrict__ dz, int N, int B, int T, int splits,
                                                       float* __restrict__ part) {
  const int sp = blockIdx.x, i = blockIdx.y, n = threadIdx.x;
  const long R = (long)T * B;
  const long per = ((R + splits - 1) / splits + 3) / 4 * 4;   // multiple of 4 so 4-row groups never straddle a time step
  const long r0 = sp * per, r1 = min(R, r0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  long r = r0;
  for (; r + 3 < r1; r += 4) {
    const long t = r / B, b = r - t * B;              // B % 4 == 0: the 4 rows share t
    const float* p = dz + (((size_t)t * N + i) * B + b) * NG + n;
    s0 += p[0]; s1 += p[NG]; s2 += p[2 * NG]; s3 += p[3 * NG];
  }
  for (; r < r1; ++r) {
    const long t = r / B, b = r - t * B;
    s0 += dz[(((size_t)t * N + i) * B + b) * NG + n];
  }
  part[((size_t)sp * N + i) * NG + n] = (s0 + s1) + (s2 + s3);
}
__global__ void dz_colsum_reduce_kernel(const __grid_constant__ nmarl_model m, const float* __restrict__ part, int splits,
                                        float* __restrict__ grads) {
  const int i = blockIdx.x, n = threadIdx.x;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += part[((size_t)sp * m.n_agent + i) * NG + n];
  grads[m.agent[i].o_b + n] = s;
}

}  // namespace

int nmarl_tc_wgrad_splits(int n_agent) {
  // 2 M-tiles x splits x agents CTAs: a whole number of waves of 148 SMs when possible
  int s = 37;
  while (2 * s * n_agent > 148 * 8 && s > 1) s = (s + 1) / 2;
  return s;
}

int64_t nmarl_tc_wgrad_ws_floats(const nmarl_model* m) {
  return (int64_t)nmarl_tc_wgrad_splits(m->n_agent) * m->n_agent * 257 * 256 + (int64_t)64 * m->n_agent * NG;
}

// gate wgrad on tensor cores.  ws must hold nmarl_tc_wgrad_ws_floats(m) floats.
int nmarl_tc_launch_gate_wgrad(const nmarl_model* m, int B, int T, const float* sv_sh, const float* dzT, const float* sv_dz,
                               float* ws, int* err, int* splits_out, cudaStream_t st) {
  TcWgK k{};
  k.N = m->n_agent; k.B = B; k.T = T; k.splits = nmarl_tc_wgrad_splits(m->n_agent); k.i_lda = m->s_dim + NH;
  k.A = sv_sh; k.BT = dzT; k.ws = ws; k.err = err;
  static bool configured = false;
  if (!configured) {
    NMARL_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    configured = true;
  }
  tc_wgrad_kernel<<<dim3(2 * k.splits, m->n_agent), TC_THREADS, TC_SMEM, st>>>(k);
  NMARL_LAUNCH_CHECK();
  *splits_out = k.splits;
  // bias: column sums of dz
  float* part = ws + (size_t)k.splits * m->n_agent * 257 * 256;
  dz_colsum_kernel<<<dim3(64, m->n_agent), 256, 0, st>>>(sv_dz, m->n_agent, B, T, 64, part);
  NMARL_LAUNCH_CHECK();
  return 0;
}

int nmarl_tc_launch_bias_reduce(const nmarl_model* m, const float* ws, int splits, float* grads, cudaStream_t st) {
  const float* part = ws + (size_t)splits * m->n_agent * 257 * 256;
  dz_colsum_reduce_kernel<<<m->n_agent, 256, 0, st>>>(*m, part, 64, grads);
  NMARL_LAUNCH_CHECK();
  return 0;
}
