// tc_gemm.cu -- weight packing for the tcgen05 path + a stand-alone 3xTF32 GEMM used by the tests to
// validate the tensor-core pipeline (TMEM alloc, tcgen05.st/ld, UMMA descriptors, bulk copies, mbarriers)
// in isolation:  C[M x N] = A[M x K] * W[K x N]   (M % 128 == 0, K % 8 == 0, N in {64, 256}).
#include "common.cuh"
#include "tc.cuh"

namespace {

// Packed operand: for every 32-wide k-block kb: [hi tile | lo tile], each tile = N rows x 128 B in the
// SWIZZLE_128B K-major layout (row n <-> output column n of W, i.e. the tile holds W^T).
__global__ void pack_b_kernel(const float* __restrict__ W, int ldw, int K, int n0, int nrows, float* __restrict__ out) {
  const int nkb = (K + 31) / 32;
  const int total = nkb * nrows * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % nrows;                 // consecutive threads -> consecutive columns of W (coalesced)
    const int kk = (idx / nrows) % 32;
    const int kb = idx / (nrows * 32);
    const int k = kb * 32 + kk;
    const float x = (k < K) ? W[(size_t)k * ldw + n0 + n] : 0.f;
    float hi, lo;
    tc::split_tf32(x, hi, lo);
    char* tile = reinterpret_cast<char*>(out) + (size_t)kb * 2 * nrows * 128;
    const uint32_t off = tc::sw128_offset(n, kk);
    *reinterpret_cast<float*>(tile + off) = hi;
    *reinterpret_cast<float*>(tile + (size_t)nrows * 128 + off) = lo;
  }
}

// Probe variant of the packing (tools/probe_tf32_operand.py): the "hi" tile holds the RAW fp32 values, the "lo" tile
// x - trunc_tf32(x) as before.  If the TF32 datapath ignores the 13 low mantissa bits of a shared-memory operand,
// a GEMM on these tiles is bit-identical to one on the masked tiles, and operand tiles written by the backward
// kernel would not need a separate hi copy.
__global__ void pack_b_raw_kernel(const float* __restrict__ W, int ldw, int K, int n0, int nrows, float* __restrict__ out) {
  const int nkb = (K + 31) / 32;
  const int total = nkb * nrows * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % nrows, kk = (idx / nrows) % 32, kb = idx / (nrows * 32);
    const int k = kb * 32 + kk;
    const float x = (k < K) ? W[(size_t)k * ldw + n0 + n] : 0.f;
    char* tile = reinterpret_cast<char*>(out) + (size_t)kb * 2 * nrows * 128;
    const uint32_t off = tc::sw128_offset(n, kk);
    *reinterpret_cast<float*>(tile + off) = x;
    *reinterpret_cast<float*>(tile + (size_t)nrows * 128 + off) = tc::tf32_lo_of_raw(x);
  }
}

template <int N>
__global__ void __launch_bounds__(192, 1) tc_gemm_test_kernel(const float* __restrict__ A, int K, const float* __restrict__ Bp,
                                                              float* __restrict__ C, int* err) {
  constexpr int S = 2;                                   // B stages
  constexpr uint32_t TILE = N * 128;                     // bytes of one hi (or lo) tile
  constexpr uint32_t STAGE = 2 * TILE;
  constexpr uint32_t ACC_COL = 0, A_COL = 256;           // TMEM columns: accumulator, then 2 x (hi32 | lo32)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bst = smem;                                   // S stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * STAGE);
  uint64_t* b_full = bars, *b_empty = bars + S, *a_full = bars + 2 * S, *a_empty = bars + 2 * S + 2, *acc_full = bars + 2 * S + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (K + 31) / 32;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&a_full[s], 128); tc::mbar_init(&a_empty[s], 1); }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ---- row threads: produce the A operand chunks, then the epilogue ----
    const size_t row = (size_t)blockIdx.x * 128 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int kb = 0; kb < nkb; ++kb) {
      const int slot = kb & 1;
      tc::mbar_wait(&a_empty[slot], ((kb >> 1) & 1) ^ 1, err, 1);
      tc::fence_after_sync();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float x[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = kb * 32 + h * 16 + q * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < K) v = *reinterpret_cast<const float4*>(A + row * K + k);
          x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
        const uint32_t col = A_COL + slot * 64 + h * 16;
        tc::tmem_st_hilo16(tmem + lane_base + col, tmem + lane_base + col + 32, x);
      }
      tc::wait_st();
      tc::fence_before_sync();
      tc::mbar_arrive(&a_full[slot]);
    }
    tc::mbar_wait(acc_full, 0, err, 2);
    tc::fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < N; c += 16) {
      float v[16];
      tc::tmem_ld16(tmem + lane_base + ACC_COL + c, v);
      tc::wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(C + row * N + c + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    tc::fence_before_sync();
  } else if (warp == 4) {
    // ---- B producer: one bulk copy (hi | lo tiles) per k-block ----
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % S;
        tc::mbar_wait(&b_empty[st], ((kb / S) & 1) ^ 1, err, 3);
        tc::mbar_arrive_expect_tx(&b_full[st], STAGE);
        tc::bulk_g2s(bst + st * STAGE, reinterpret_cast<const uint8_t*>(Bp) + (size_t)kb * STAGE, STAGE, &b_full[st]);
      }
    }
  } else {
    // ---- MMA issuer ----
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(128, N);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % S, slot = kb & 1;
        tc::mbar_wait(&b_full[st], (kb / S) & 1, err, 4);
        tc::mbar_wait(&a_full[slot], (kb >> 1) & 1, err, 5);
        tc::fence_after_sync();
        const int ksteps = min(4, (K - kb * 32) / 8);
        const uint64_t d_hi = tc::smem_desc_sw128(bst + st * STAGE), d_lo = tc::smem_desc_sw128(bst + st * STAGE + TILE);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t a_hi = tmem + A_COL + slot * 64 + ks * 8, a_lo = a_hi + 32;
          tc::mma_tf32_ts(tmem + ACC_COL, a_hi, d_hi + 2 * ks, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          tc::mma_tf32_ts(tmem + ACC_COL, a_hi, d_lo + 2 * ks, idesc, 1u);
          tc::mma_tf32_ts(tmem + ACC_COL, a_lo, d_hi + 2 * ks, idesc, 1u);
        }
        tc::mma_commit(&a_empty[slot]);
        tc::mma_commit(&b_empty[st]);
      }
      tc::mma_commit(acc_full);
    }
  }
  __syncthreads();
  if (warp == 5) { tc::fence_after_sync(); tc::tmem_dealloc(tmem, 512); }
}

// Probe for DESIGN.md 6.1: the same GEMM with the A operand staged in 128B-swizzled SHARED memory (both tcgen05.mma
// operands from shared-memory descriptors) instead of TMEM -- the form a two-tiles-per-SM cell kernel needs, since
// it leaves only the 256 accumulator columns of a tile in TMEM.  Not used by the product path.
template <int N>
__global__ void __launch_bounds__(192, 1) tc_gemm_ss_test_kernel(const float* __restrict__ A, int K, const float* __restrict__ Bp,
                                                                 float* __restrict__ C, int* err) {
  constexpr int S = 2;
  constexpr uint32_t TILE = N * 128, STAGE = 2 * TILE;
  constexpr uint32_t A_TILE = 128 * 128, A_SLOT = 2 * A_TILE;          // [hi | lo], 128 rows x 32 floats each
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bst = smem;
  uint8_t* ast = smem + S * STAGE;                                      // 2 A slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(ast + 2 * A_SLOT);
  uint64_t* b_full = bars, *b_empty = bars + S, *a_full = bars + 2 * S, *a_empty = bars + 2 * S + 2, *acc_full = bars + 2 * S + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (K + 31) / 32;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&a_full[s], 128); tc::mbar_init(&a_empty[s], 1); }
    tc::mbar_init(acc_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, 256);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    const size_t row = (size_t)blockIdx.x * 128 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int kb = 0; kb < nkb; ++kb) {
      const int slot = kb & 1;
      tc::mbar_wait(&a_empty[slot], ((kb >> 1) & 1) ^ 1, err, 1);
      uint8_t* hi_t = ast + slot * A_SLOT, *lo_t = hi_t + A_TILE;
#pragma unroll
      for (int q = 0; q < 8; ++q) {                                       // eight 16-byte chunks of this thread's row
        const int k = kb * 32 + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) v = *reinterpret_cast<const float4*>(A + row * K + k);
        float4 h, l;
        tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
        tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
        const uint32_t off = tc::sw128_offset((uint32_t)tid, (uint32_t)(q * 4));
        *reinterpret_cast<float4*>(hi_t + off) = h;
        *reinterpret_cast<float4*>(lo_t + off) = l;
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&a_full[slot]);
    }
    tc::mbar_wait(acc_full, 0, err, 2);
    tc::fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < N; c += 16) {
      float v[16];
      tc::tmem_ld16(tmem + lane_base + c, v);
      tc::wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(C + row * N + c + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    tc::fence_before_sync();
  } else if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % S;
        tc::mbar_wait(&b_empty[st], ((kb / S) & 1) ^ 1, err, 3);
        tc::mbar_arrive_expect_tx(&b_full[st], STAGE);
        tc::bulk_g2s(bst + st * STAGE, reinterpret_cast<const uint8_t*>(Bp) + (size_t)kb * STAGE, STAGE, &b_full[st]);
      }
    }
  } else {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(128, N);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb % S, slot = kb & 1;
        tc::mbar_wait(&b_full[st], (kb / S) & 1, err, 4);
        tc::mbar_wait(&a_full[slot], (kb >> 1) & 1, err, 5);
        tc::fence_after_sync();
        const int ksteps = min(4, (K - kb * 32) / 8);
        const uint64_t b_hi = tc::smem_desc_sw128(bst + st * STAGE), b_lo = tc::smem_desc_sw128(bst + st * STAGE + TILE);
        const uint64_t a_hi = tc::smem_desc_sw128(ast + slot * A_SLOT), a_lo = tc::smem_desc_sw128(ast + slot * A_SLOT + A_TILE);
        for (int ks = 0; ks < ksteps; ++ks) {
          tc::mma_tf32_ss(tmem, a_hi + 2 * ks, b_hi + 2 * ks, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          tc::mma_tf32_ss(tmem, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
          tc::mma_tf32_ss(tmem, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
        }
        tc::mma_commit(&a_empty[slot]);
        tc::mma_commit(&b_empty[st]);
      }
      tc::mma_commit(acc_full);
    }
  }
  __syncthreads();
  if (warp == 5) { tc::fence_after_sync(); tc::tmem_dealloc(tmem, 256); }
}

}  // namespace

int nmarl_launch_pack_b(const float* W, int ldw, int K, int n0, int nrows, float* out, cudaStream_t st) {
  const int total = ((K + 31) / 32) * nrows * 32;
  pack_b_kernel<<<(total + 255) / 256, 256, 0, st>>>(W, ldw, K, n0, nrows, out);
  NMARL_LAUNCH_CHECK();
  return 0;
}

static int tc_gemm_selftest_impl(const float* A, const float* W, float* C, int M, int K, int N, float* scratch, int* err,
                                 void* stream, bool raw_hi);

// C[M x N] = A[M x K] * W[K x N]; scratch must hold ceil(K/32) * 2 * N * 32 floats; err: device int (0 = ok)
extern "C" __attribute__((visibility("default"))) int nmarl_tc_gemm_selftest(const float* A, const float* W, float* C, int M, int K,
                                                                               int N, float* scratch, int* err, void* stream) {
  return tc_gemm_selftest_impl(A, W, C, M, K, N, scratch, err, stream, false);
}

// same GEMM with the un-masked operand in the hi tile (hardware probe, not used by the product path)
extern "C" __attribute__((visibility("default"))) int nmarl_tc_gemm_selftest_raw(const float* A, const float* W, float* C, int M,
                                                                                   int K, int N, float* scratch, int* err,
                                                                                   void* stream) {
  return tc_gemm_selftest_impl(A, W, C, M, K, N, scratch, err, stream, true);
}

static int tc_gemm_selftest_impl(const float* A, const float* W, float* C, int M, int K, int N, float* scratch, int* err,
                                 void* stream, bool raw_hi) {
  NMARL_CHECK(M > 0 && M % 128 == 0 && K > 0 && K % 8 == 0 && (N == 64 || N == 256), "tc_gemm_selftest: unsupported shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (raw_hi) {
    const int total = ((K + 31) / 32) * N * 32;
    pack_b_raw_kernel<<<(total + 255) / 256, 256, 0, st>>>(W, N, K, 0, N, scratch);
    NMARL_LAUNCH_CHECK();
  } else if (nmarl_launch_pack_b(W, N, K, 0, N, scratch, st)) return 1;
  const size_t smem = 2 * 2 * (size_t)N * 128 + 1024 + 256;
  if (N == 256) {
    NMARL_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_test_kernel<256><<<M / 128, 192, smem, st>>>(A, K, scratch, C, err);
  } else {
    NMARL_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_test_kernel<64><<<M / 128, 192, smem, st>>>(A, K, scratch, C, err);
  }
  NMARL_LAUNCH_CHECK();
  return 0;
}

// same GEMM with both operands in shared memory (hardware probe for the two-tiles-per-SM plan; not on the product path)
extern "C" __attribute__((visibility("default"))) int nmarl_tc_gemm_selftest_ss(const float* A, const float* W, float* C, int M,
                                                                                  int K, int N, float* scratch, int* err,
                                                                                  void* stream) {
  NMARL_CHECK(M > 0 && M % 128 == 0 && K > 0 && K % 8 == 0 && (N == 64 || N == 256), "tc_gemm_selftest_ss: unsupported shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (nmarl_launch_pack_b(W, N, K, 0, N, scratch, st)) return 1;
  const size_t smem = 2 * 2 * (size_t)N * 128 + 2 * 2 * 128 * 128 + 1024 + 256;
  if (N == 256) {
    NMARL_CUDA(cudaFuncSetAttribute(tc_gemm_ss_test_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_ss_test_kernel<256><<<M / 128, 192, smem, st>>>(A, K, scratch, C, err);
  } else {
    NMARL_CUDA(cudaFuncSetAttribute(tc_gemm_ss_test_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_ss_test_kernel<64><<<M / 128, 192, smem, st>>>(A, K, scratch, C, err);
  }
  NMARL_LAUNCH_CHECK();
  return 0;
}
