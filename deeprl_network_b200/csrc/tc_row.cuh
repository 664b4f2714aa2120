// tc_row.cuh -- shared structure of the tcgen05 cell kernels (forward: tc_cell.cu, backward: tc_bwd.cu):
// TMEM column map, k-block schedule entries, the row-thread producer helpers and the B-producer / MMA-issuer
// role loops.  See tc_cell.cu for the overall design.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace tcrow {

constexpr int S_STAGES = 3;
constexpr uint32_t STAGE_BYTES = 2 * 256 * 128;          // hi+lo tiles of the widest operand (N = 256)
constexpr uint32_t ACC_COL = 0, A_COL = 256;            // TMEM: [0,256) accumulators (encoders reuse it), [256,512) A ring
// A-operand ring in TMEM: A_SLOTS slots of (hi 32 | lo 32) columns.  4 slots use the whole second half of TMEM; the row
// threads then run up to four k-blocks ahead of the MMA issuer, which matters for the short encoder GEMMs whose
// MMAs (N = 64) finish faster than a produce -> commit -> a_empty round trip.
constexpr int A_SLOTS = 4;
static_assert((A_SLOTS & (A_SLOTS - 1)) == 0 && A_COL + A_SLOTS * 64 <= 512, "A ring must fit TMEM");
constexpr int MAX_KB = 40;
// NSET warp-sets share every env row: set s of row r works on columns [s*W, (s+1)*W) of each 32-wide input
// k-block and on hidden units [s*EW, (s+1)*EW) of the encoders / LSTM cell.  4 sets = 16 row warps per SM
// (4 per scheduler) so global-load, TMEM and barrier latencies overlap across warps.
constexpr int NSET = 4;
constexpr int W = 32 / NSET, EW = 64 / NSET;
constexpr int ROW_THREADS = 128 * NSET;
constexpr int TC_THREADS = ROW_THREADS + 64;
static_assert(W % 8 == 0 && EW % 8 == 0, "8-column TMEM pieces");

struct KbEnt {
  uint32_t off_bytes, bytes;
  uint16_t dcol;                                          // accumulator column of this GEMM
  uint8_t ksteps, first, last_enc, last_acc, pad0, pad1;
};

struct RowCtx {
  uint32_t tmem, lane_base;
  uint64_t* a_full; uint64_t* a_empty; uint64_t* enc_full;
  int q, e, set;
  int* err;
};

__device__ __forceinline__ void produce_begin(RowCtx& c) {
  const int slot = c.q & (A_SLOTS - 1);
  tc::mbar_wait(&c.a_empty[slot], ((c.q / A_SLOTS) & 1) ^ 1, c.err, 11);
  tc::fence_after_sync();
}
__device__ __forceinline__ void produce_piece(RowCtx& c, int col /*0..31, multiple of 8*/, const float (&x)[8]) {
  const uint32_t t = c.tmem + c.lane_base + A_COL + (c.q & (A_SLOTS - 1)) * 64 + col;
  tc::tmem_st_hilo8(t, t + 32, x);
}
__device__ __forceinline__ void produce_end(RowCtx& c) {
  tc::wait_st();
  tc::fence_before_sync();
  tc::mbar_arrive(&c.a_full[c.q & (A_SLOTS - 1)]);
  c.q++;
}
// one input k-block: this thread contributes columns [set*W, set*W + W)
__device__ __forceinline__ void produce_in(RowCtx& c, const float (&x)[W]) {
  produce_begin(c);
#pragma unroll
  for (int p = 0; p < W / 8; ++p) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = x[8 * p + j];
    produce_piece(c, c.set * W + 8 * p, t);
  }
  produce_end(c);
}
// two k-blocks fed by a 64-wide activation vector of which this thread holds [set*EW, set*EW + EW)
__device__ __forceinline__ void produce_act(RowCtx& c, const float (&s)[EW]) {
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
    produce_begin(c);
#pragma unroll
    for (int p = 0; p < EW / 8; ++p) {
      const int col = c.set * EW + 8 * p;
      if ((col >> 5) == hb) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = s[8 * p + j];
        produce_piece(c, col & 31, t);
      }
    }
    produce_end(c);
  }
}
// wait for the encoder GEMMs issued so far
__device__ __forceinline__ void enc_wait(RowCtx& c) {
  tc::mbar_wait(c.enc_full, c.e & 1, c.err, 12);
  c.e++;
  tc::fence_after_sync();
}
// this thread's EW columns of a 64-wide result block at accumulator column `col`
__device__ __forceinline__ void enc_load(RowCtx& c, uint32_t col, float (&v)[EW]) {
#pragma unroll
  for (int p = 0; p < EW / 8; ++p) {
    float t[8];
    tc::tmem_ld8(c.tmem + c.lane_base + col + c.set * EW + 8 * p, t);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) v[8 * p + j] = t[j];
  }
  tc::fence_before_sync();
}
// feature-major saved activations [feature][env]: lane == env row, so one warp access per feature is a single
// 128-byte line (the row-major form would touch 32 lines per access)
template <int NV>
__device__ __forceinline__ void st_fm(float* base, int f0, int B, int b, const float (&v)[NV]) {
  // saved activations are written once and read once much later (BPTT): streaming stores keep them from
  // evicting the weights and the recurrent state the next calls re-read from L2
#pragma unroll
  for (int j = 0; j < NV; ++j) __stcs(base + (size_t)(f0 + j) * B + b, v[j]);
}
template <int NV>
__device__ __forceinline__ void ld_fm(const float* base, int f0, int B, int b, float (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = __ldcs(base + (size_t)(f0 + j) * B + b);
}
// state tensors (h, c, messages and their gradients): plane p of [planes][B][64] (env-major, FM = false) or
// [planes][64][B] (feature-major, FM = true: lane == env row -> one 128-byte line per warp access)
template <bool FM, int NV>
__device__ __forceinline__ void ld_state(const float* base, size_t plane, int b, int u0, int B, float (&v)[NV]) {
  if (FM) {
    const float* p = base + (plane * NH + u0) * (size_t)B + b;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = p[(size_t)j * B];
  } else {
    const float* p = base + (plane * (size_t)B + b) * NH + u0;
#pragma unroll
    for (int q = 0; q < NV / 4; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(p + 4 * q);
      v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
  }
}
template <bool FM, int NV>
__device__ __forceinline__ void st_state(float* base, size_t plane, int b, int u0, int B, const float (&v)[NV]) {
  if (FM) {
    float* p = base + (plane * NH + u0) * (size_t)B + b;
#pragma unroll
    for (int j = 0; j < NV; ++j) p[(size_t)j * B] = v[j];
  } else {
    float* p = base + (plane * (size_t)B + b) * NH + u0;
#pragma unroll
    for (int q = 0; q < NV / 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
}
template <int NV>
__device__ __forceinline__ void store_vec(float* dst, const float (&s)[NV]) {
#pragma unroll
  for (int q = 0; q < NV / 4; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
}
__device__ __forceinline__ void bias_act(float (&v)[EW], const float* __restrict__ b, int act /*0 relu 1 tanh 2 none*/) {
#pragma unroll
  for (int q = 0; q < EW / 4; ++q) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + q);
    const float z[4] = {v[4 * q] + bb.x, v[4 * q + 1] + bb.y, v[4 * q + 2] + bb.z, v[4 * q + 3] + bb.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) v[4 * q + j] = act == 0 ? fmaxf(z[j], 0.f) : (act == 1 ? tanhf(z[j]) : z[j]);
  }
}
// MUFU-based activations for the tensor-core epilogues (ex2.approx + rcp): absolute error ~1e-7, well inside the
// 1e-5 parity budget, ~6x fewer instructions than expf/tanhf + IEEE division.
__device__ __forceinline__ float frcp_(float x) {          // one MUFU.RCP (1 ulp), no IEEE fix-up path
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fsigmoid(float x) { return frcp_(1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return fmaf(2.0f, frcp_(1.0f + __expf(-2.0f * x)), -1.0f); }
__device__ __forceinline__ void row_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory"); }


// ---- role loops (one elected thread each) ------------------------------------------------------------------
__device__ __forceinline__ void producer_loop(const KbEnt* sched, int n_kb, uint8_t* bst, uint64_t* b_full, uint64_t* b_empty,
                                              const float* wpack, int* err) {
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(wpack);
  for (int q = 0; q < n_kb; ++q) {
    const int st = q % S_STAGES;
    tc::mbar_wait(&b_empty[st], ((q / S_STAGES) & 1) ^ 1, err, 21);
    const KbEnt e = sched[q];
    tc::mbar_arrive_expect_tx(&b_full[st], e.bytes);
    tc::bulk_g2s(bst + st * STAGE_BYTES, wp + e.off_bytes, e.bytes, &b_full[st]);
  }
}
__device__ __forceinline__ void mma_loop(const KbEnt* sched, int n_kb, uint8_t* bst, uint64_t* b_full, uint64_t* b_empty,
                                         uint64_t* a_full, uint64_t* a_empty, uint64_t* enc_full, uint64_t* acc_full,
                                         uint32_t tmem, int* err, long long* prof = nullptr) {
  for (int q = 0; q < n_kb; ++q) {
    const int st = q % S_STAGES, slot = q & (A_SLOTS - 1);
    const KbEnt e = sched[q];
    tc::mbar_wait(&b_full[st], (q / S_STAGES) & 1, err, 31);
    if (prof) prof[32 + 3 * q] = clock64();
    tc::mbar_wait(&a_full[slot], (q / A_SLOTS) & 1, err, 32);
    tc::fence_after_sync();
    if (prof) prof[33 + 3 * q] = clock64();
    const uint32_t tile = e.bytes / 2;
    const uint32_t ncols = tile / 128;                       // N of this operand
    const uint64_t d_hi = tc::smem_desc_sw128(bst + st * STAGE_BYTES), d_lo = tc::smem_desc_sw128(bst + st * STAGE_BYTES + tile);
    const uint32_t idesc = tc::idesc_tf32(128, ncols);
    const uint32_t dcol = tmem + e.dcol;
    for (int ks = 0; ks < e.ksteps; ++ks) {
      const uint32_t a_hi = tmem + A_COL + slot * 64 + ks * 8, a_lo = a_hi + 32;
      tc::mma_tf32_ts(dcol, a_hi, d_hi + 2 * ks, idesc, (e.first && ks == 0) ? 0u : 1u);
      tc::mma_tf32_ts(dcol, a_hi, d_lo + 2 * ks, idesc, 1u);
      tc::mma_tf32_ts(dcol, a_lo, d_hi + 2 * ks, idesc, 1u);
    }
    tc::mma_commit(&a_empty[slot]);
    tc::mma_commit(&b_empty[st]);
    if (e.last_enc) tc::mma_commit(enc_full);
    if (e.last_acc) tc::mma_commit(acc_full);
    if (prof) prof[34 + 3 * q] = clock64();
  }
}
__device__ __forceinline__ KbEnt make_kb(int off_floats, int N, int K, int kb, int dcol, int first, int last_enc, int last_acc) {
  KbEnt e;
  e.off_bytes = (uint32_t)(off_floats + kb * 2 * N * 32) * 4u;
  e.bytes = 2u * N * 128u;
  const int k8 = (K + 7) / 8 * 8;
  e.ksteps = (uint8_t)min(4, (k8 - kb * 32) / 8);
  e.dcol = (uint16_t)dcol; e.first = first; e.last_enc = last_enc; e.last_acc = last_acc;
  e.pad0 = e.pad1 = 0;
  return e;
}
constexpr size_t TC_SMEM = S_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 32 * 8 /*mbarriers*/ + 16 + MAX_KB * sizeof(KbEnt) + NSET * 128 * 8 * sizeof(float);

}  // namespace tcrow
