// tc.cuh -- Blackwell (sm_100a) tensor-core primitives used by the 3xTF32 GEMMs:
// tcgen05.mma kind::tf32 with the A operand in TMEM, the B operand in 128B-swizzled shared memory
// (staged by cp.async.bulk + mbarrier), FP32 accumulators in TMEM, tcgen05.ld/st for the
// register <-> TMEM traffic.  Inline PTX only (no CUTLASS dependency).
//
// fp32-accurate products on TF32 tensor cores ("3xTF32"): x = hi + lo with hi = x ROUNDED to TF32
// (cvt.rna.tf32) and lo = the remainder x - hi (exact in fp32) rounded to TF32 as well.
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi; the dropped terms are O(2^-22 |a||b|) and -- because both parts are
// rounded to nearest -- zero-mean.  (The cheaper split, hi = x & 0xFFFFE000 with the hardware truncating lo, makes
// hi, lo and the dropped lo*lo term all err toward zero: a systematic ~2e-7 relative bias per product that does not
// average out over long contractions and is amplified by cancellation -- measured 3.6e-5 of max|g| on the 245 760-row
// weight gradients at B = 4096, T = 60, 60x the error of plain fp32 arithmetic.)  Accumulation is FP32 in TMEM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

// TF32 rounding / 3xTF32 operand split (see the header comment)
__device__ __forceinline__ uint32_t tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rn(x);
  lo = tf32_rn(x - __uint_as_float(hi));
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  split_tf32(x, h, l);
  hi = __uint_as_float(h); lo = __uint_as_float(l);
}
// remainder of a RAW fp32 operand the tensor core reads truncated (it ignores the 13 low mantissa bits)
__device__ __forceinline__ float tf32_lo_of_raw(float x) {
  return __uint_as_float(tf32_rn(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the
// stream is still running; it must execute pdl_wait() before touching anything the predecessor wrote (the wait returns
// once the predecessor grid has completed and its writes are visible).  pdl_launch_dependents() lets the NEXT
// kernel's CTAs be scheduled onto free SMs as soon as every CTA of this grid has started.  Both are no-ops for a
// kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// One elected lane of a converged warp.  ptxas treats a branch on elect.sync as single-threaded, so the operands of the
// tcgen05.mma / commit instructions inside it move to uniform registers with a plain R2UR instead of the per-operand
// ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop it emits under `if (lane == 0)` (~100 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU.  The deadline is in SM clock cycles (~70 ms); once any
// wait of the grid has timed out (*err != 0) every later wait gives up immediately so the kernel drains.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
    // The watchdog (clock + the grid-wide abort flag in global memory) is consulted only every 64th failed try_wait:
    // a global load per poll keeps the thread away from the barrier for an L2 round trip (~700 cycles) and that
    // latency was added to every wake-up of every fine-grained pipeline wait.
    if ((spins & 63u) != 63u) continue;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    if (now - t0 > (1ll << 27)) break;
    if (err != nullptr && *reinterpret_cast<volatile int*>(err) != 0) return false;
  }
  if (err != nullptr) atomicCAS(err, 0, code);
  return false;
}

// ---- bulk async copy global -> shared (TMA engine, 1-D) -------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, 16 consecutive columns: thread i of the warp <-> TMEM lane (warp%4)*32 + i
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_hilo8(uint32_t taddr_hi, uint32_t taddr_lo, const float (&x)[8]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    split_tf32(x[i], hi[i], lo[i]);
  }
  tmem_st8(taddr_hi, hi);
  tmem_st8(taddr_lo, lo);
}

// hi/lo split of 16 fp32 values and store as two TF32 operand chunks (hi at col, lo at col + lo_off)
__device__ __forceinline__ void tmem_st_hilo16(uint32_t taddr_hi, uint32_t taddr_lo, const float (&x)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    split_tf32(x[i], hi[i], lo[i]);
  }
  tmem_st16(taddr_hi, hi);
  tmem_st16(taddr_lo, lo);
}

// ---- UMMA descriptors --------------------------------------------------------------------------------------
// K-major operand tile [rows][32 tf32] in SWIZZLE_128B layout: row r at (r/8)*1024 + (r%8)*128 bytes, the
// eight 16-byte chunks of a row XOR-ed with (r%8).  SBO = 1024 B (8-row groups), LBO unused, version 1.
__device__ __forceinline__ uint64_t smem_desc_sw128(const void* tile) {
  const uint64_t addr = (uint64_t)(smem_u32(tile) >> 4) & 0x3FFFull;
  return addr | (0ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=F32, A=B=TF32, both K-major, shape M x N
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// byte offset of element (row, k) inside a swizzled [rows][32] tile
__host__ __device__ inline uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 2) ^ row) & 7u) << 4) + ((k & 3u) << 2);
}

// D[tmem_d] (+)= A[tmem_a : 128 lanes x 8 cols tf32] * B[smem desc : N rows x 8 tf32]^T   (one elected thread)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with BOTH operands in shared memory (A: 128 rows x 8 tf32 K-major tile, same descriptor format as B)
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// operand tiles written with ordinary st.shared must be made visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace tc
