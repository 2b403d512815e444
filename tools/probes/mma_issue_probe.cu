// mma_issue_probe.cu -- what does ONE tcgen05.mma cost the issuing warp, and does the issue style matter?
//   style 0: `if (lane == 0) { loop }`  -- divergent code: operands live in vector registers, the compiler wraps every
//            tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY "waterfall" (~16 SASS instructions per MMA)
//   style 1: the whole warp runs the loop, only the instruction itself is predicated by elect.sync; every operand is
//            derived from kernel parameters / __shfl_sync(...,0) so the compiler keeps it in uniform registers
// Reports cycles per MMA for the issue loop alone and until the commit lands (16 MMAs per "tile", K advance = +32 B).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gansformer-reproducibility-challenge_b200/csrc/gf_tc_common.cuh"
using namespace gf::tc;

__host__ __device__ constexpr uint32_t idesc(int M, int N, int bmn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma_ss_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b32 r;\n\t"
      "elect.sync r|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b32 r;\n\t"
      "elect.sync r|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}

template <int STYLE, int M, int N>
__global__ void probe(int tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base, 0);
  const uint32_t sa = smem_u32(smem), sb = sa + 64 * 1024;
  constexpr uint32_t ID = idesc(M, N, 0);
  if (warp == 0) {
    const uint64_t da = mkdesc(sa, 16, 1024, 2), db = mkdesc(sb, 16, 1024, 2);
    if (STYLE == 0) {
      if ((threadIdx.x & 31) == 0) {
        for (int rep = 0; rep < 2; ++rep) {
          const long long t0 = clock64();
          for (int t = 0; t < tiles; ++t) {
            const uint32_t d = tmem + 256 + (t & 1) * 64;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) umma_ss(d, da + (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2), db + (uint64_t)((kk >> 2) * 256 + (kk & 3) * 2), ID, kk ? 1u : 0u);
          }
          const long long t1 = clock64();
          umma_commit(smem_u32(&bar));
          mbar_wait(smem_u32(&bar), (uint32_t)rep);
          const long long t2 = clock64();
          if (rep == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
      }
    } else {
      for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
          const uint32_t d = tmem + 256 + (t & 1) * 64;
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) umma_ss_elect(d, da + (uint64_t)((kk >> 2) * 1024 + (kk & 3) * 2), db + (uint64_t)((kk >> 2) * 256 + (kk & 3) * 2), ID, kk ? 1u : 0u);
        }
        const long long t1 = clock64();
        umma_commit_elect(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), (uint32_t)rep);
        const long long t2 = clock64();
        if (rep == 1 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory"); }
}

namespace gf { void set_error(const char*, ...) {} void set_path(int) {} void note_launch() {} void set_centroid_path(int) {} }

template <int STYLE, int M, int N>
static void run(long long* o) {
  const int tiles = 64;
  long long ho[2];
  cudaFuncSetAttribute(probe<STYLE, M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<STYLE, M, N><<<1, 128, 200 * 1024>>>(tiles, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
  printf("style %d  M=%3d N=%3d | issue %7.1f cyc/MMA   done %7.1f cyc/MMA\n", STYLE, M, N, (double)ho[0] / (tiles * 16), (double)ho[1] / (tiles * 16));
}

int main() {
  long long* o; cudaMalloc(&o, 16);
  run<0, 128, 16>(o); run<1, 128, 16>(o);
  run<0, 128, 32>(o); run<1, 128, 32>(o);
  run<0, 64, 32>(o);  run<1, 64, 32>(o);
  run<0, 64, 128>(o); run<1, 64, 128>(o);
  run<0, 128, 256>(o); run<1, 128, 256>(o);
  return 0;
}
