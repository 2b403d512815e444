// gf_tc_gemm.cu -- small TF32 tensor-core GEMM for the per-image products between the big kernels of a layer:
//     C[M,N] = alpha * A[M,K] . B[K,N] + E[(row % emod), :] + v[:]          (A, B row-major, lda = K, ldb = N)
// used (TF32 mode only) for KPALL = keys . AK + CK, MALL = Y . AM + CM and the duplex centroids Xbar . Wv2 + bv2
// (reference side, expected src/training/network.py: the K / V dense_layer calls of transformer_layer).  M = B*k is a few
// thousand rows, so the CUDA-core SGEMM (gf_fold.cu) spent ~50 us per product at C = 512 -- a third of a small duplex
// layer.  One 128 x 64 output tile per CTA, BK = 32:
//   warp 0  TMA producer: A tile [128 x 32] SWIZZLE_128B (K-major operand), B tile [32 x 64] as two [32 x 32] boxes with
//           SWIZZLE_128B_ATOM_32B (B is row-major [K,N] = MN-major for the tensor core; see gf_tc_cen.cu)
//   warp 1  MMA issuer, warp-converged (uniform-register descriptors): 4 tcgen05.mma (M=128, N=64, K=8) per stage
//   warps 2-5  epilogue: TMEM -> registers -> alpha, bias rows -> global (thread = output row)
// Out-of-range rows / columns / k are zero-filled by TMA and masked at the store.  Operands are truncated to TF32 by the
// tensor core; alpha carries the mean-truncation compensation of both operands (gf_fold.cu: GF_TF32_TRUNC_COMP).
#include <stdlib.h>
#include "gf_common.cuh"
#include "gf_tc_common.cuh"

namespace gf {
namespace tcg {

using namespace tc;

constexpr int BM = 128, BN = 64, BK = 32;
constexpr int A_BYTES = BM * BK * 4;               // 16 KB
constexpr int B_BYTES = BK * BN * 4;               // 8 KB: two 4 KB blocks of 32 n
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NSTAGES = 6;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 64;

struct Bars {
  uint64_t full[NSTAGES], empty[NSTAGES];
  uint64_t acc_full;
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                          // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ Cm, int ldc,
               int M, int N, int K, float alpha, const float* __restrict__ E, int lde, int emod, const float* __restrict__ v) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  const uint32_t s_base = smem_u32(smem_raw) + pad;
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + pad + NSTAGES * STAGE_BYTES);
  const uint32_t s_bars = s_base + NSTAGES * STAGE_BYTES;
  auto bar = [&](const void* p) -> uint32_t { return s_bars + (uint32_t)((const uint8_t*)p - (const uint8_t*)bars); };
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int nk = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int i = 0; i < NSTAGES; ++i) { mbar_init(bar(&bars->full[i]), 1); mbar_init(bar(&bars->empty[i]), 1); }
    mbar_init(bar(&bars->acc_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar(&bars->tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t ph = 0;
      for (int kt = 0; kt < nk; ++kt) {
        mbar_wait(bar(&bars->empty[stage]), ph ^ 1u);
        const uint32_t fb = bar(&bars->full[stage]);
        const uint32_t sa = s_base + stage * STAGE_BYTES;
        mbar_expect_tx(fb, STAGE_BYTES);
        tma_load_2d(sa, &tmA, fb, kt * BK, m0);
        tma_load_2d(sa + A_BYTES, &tmB, fb, n0, kt * BK);
        tma_load_2d(sa + A_BYTES + 4096, &tmB, fb, n0 + 32, kt * BK);
        if (++stage == NSTAGES) { stage = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: tf32 x tf32 -> f32, A K-major, B MN-major, N = 64, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t dA0 = umma_desc(s_base, 1024, LAYOUT_SW128);
    const uint64_t dB0 = desc_mn(s_base + A_BYTES, 4096, 512);
    int stage = 0; uint32_t ph = 0;
    for (int kt = 0; kt < nk; ++kt) {
      mbar_wait(bar(&bars->full[stage]), ph);
      tc_fence_after();
      const uint64_t da = dA0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
      const uint64_t db = dB0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_ss_elect(tmem, da + kk * 2, db + (uint64_t)(kk * 64), IDESC, (kt | kk) ? 1u : 0u);
      umma_commit_elect(bar(&bars->empty[stage]));
      if (++stage == NSTAGES) { stage = 0; ph ^= 1u; }
    }
    umma_commit_elect(bar(&bars->acc_full));
  } else {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    mbar_wait(bar(&bars->acc_full), 0);
    tc_fence_after();
    const float* erow = E ? E + (size_t)(row % emod) * lde : nullptr;
    float* crow = Cm + (size_t)row * ldc;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float acc[16];
      tmem_ld16(tmem + lane_addr + c0, acc);
      tmem_wait_ld();
      if (row < M) {
        const int cb = n0 + c0;
        if (cb + 16 <= N && (ldc & 3) == 0) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            float4 r;
            r.x = alpha * acc[i4 * 4 + 0]; r.y = alpha * acc[i4 * 4 + 1]; r.z = alpha * acc[i4 * 4 + 2]; r.w = alpha * acc[i4 * 4 + 3];
            const int c = cb + i4 * 4;
            if (erow) { r.x += erow[c]; r.y += erow[c + 1]; r.z += erow[c + 2]; r.w += erow[c + 3]; }
            if (v) { r.x += v[c]; r.y += v[c + 1]; r.z += v[c + 2]; r.w += v[c + 3]; }
            *reinterpret_cast<float4*>(crow + c) = r;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = cb + i;
            if (c < N) crow[c] = alpha * acc[i] + (erow ? erow[c] : 0.f) + (v ? v[c] : 0.f);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace tcg

// lda == K and ldb == N (dense row-major operands); returns GF_ERR_UNSUPPORTED when the shape cannot be described to TMA
bool gemm_tc_ok(int M, int N, int K, const float* A, const float* B, const float* Cm, int ldc) {
  static const bool disabled = getenv("GF_DISABLE_TC") != nullptr || getenv("GF_DISABLE_TC_GEMM") != nullptr;
  if (disabled || M < 1 || N < 32 || K < 4) return false;
  if ((K & 3) || (N & 3)) return false;                                   // TMA: row strides are multiples of 16 bytes
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)Cm & 15) || (ldc & 3)) return false;
  return true;
}

int gemm_tc(cudaStream_t st, int M, int N, int K, const float* A, const float* B, float* Cm, int ldc, float alpha,
            const float* E, int lde, int emod, const float* v) {
  using namespace tcg;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = tc::make_map(&tmA, A, (uint64_t)M, (uint64_t)K, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = tc::make_map(&tmB, B, (uint64_t)K, (uint64_t)N, BK, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  const int smem_bytes = NSTAGES * STAGE_BYTES + (int)sizeof(Bars) + 1024;
  GF_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));   // per device: set on every call
  if (emod < 1) emod = 1;
  // both operands are truncated to TF32 by the tensor core: compensate the mean truncation bias of each (0.7213 * 2^-11)
  const float comp = 1.000352220f * 1.000352220f;
  gemm_tc_kernel<<<dim3((N + BN - 1) / BN, (M + BM - 1) / BM), NUM_THREADS, smem_bytes, st>>>(tmA, tmB, Cm, ldc, M, N, K, alpha * comp, E, lde, emod, v);
  GF_LAUNCH_OK();
  return GF_OK;
}

}  // namespace gf
