// mma_probe.cu -- cost model of small tcgen05.mma (kind::tf32) instructions on sm_100a.
// One CTA; one thread issues REPS back-to-back MMAs of a given shape / operand layout / accumulator pattern, commits, and
// the elapsed SM cycles (clock64) until the commit's mbarrier completes are reported per MMA.  Operand contents are zero.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gansformer-reproducibility-challenge_b200/csrc/gf_tc_common.cuh"
using namespace gf::tc;

constexpr uint32_t L_SW128 = 2, L_BASE32B = 1;
__host__ __device__ constexpr uint32_t idesc(int M, int N, int bmn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)layout << 61;
  return d;
}
struct Case { int M, N, bmn, ndst, ts, reps; };

__global__ void probe(const Case* cases, int ncases, long long* out) {
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
  const uint32_t tmem = tmem_base, sa = smem_u32(smem), sb = sa + 64 * 1024;
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    for (int c = 0; c < ncases; ++c) {
      const Case cs = cases[c];
      const uint32_t id = idesc(cs.M, cs.N, cs.bmn);
      const uint64_t da = desc(sa, 16, 1024, L_SW128);
      const uint64_t db = cs.bmn ? desc(sb, 16384, 512, L_BASE32B) : desc(sb, 16, 1024, L_SW128);
      for (int rep = 0; rep < 2; ++rep) {             // rep 0 warms up
        const long long t0 = clock64();
        for (int i = 0; i < cs.reps; ++i) {
          const uint32_t d = tmem + 256 + (uint32_t)((i % cs.ndst) * 32) % 256;
          if (cs.ts) umma_ts(d, tmem + (i & 3) * 8, db + (uint64_t)((i & 3) * 2), id, 1u);
          else umma_ss(d, da + (uint64_t)((i & 3) * 2), cs.bmn ? db + (uint64_t)((i & 15) * 64) : db + (uint64_t)((i & 3) * 2), id, 1u);
        }
        const long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase); phase ^= 1;
        const long long t2 = clock64();
        if (rep == 1) { out[2 * c] = t1 - t0; out[2 * c + 1] = t2 - t0; }
      }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory"); }
}

namespace gf { void set_error(const char*, ...) {} void set_path(int) {} void note_launch() {} void set_centroid_path(int) {} }

int main() {
  const Case h[] = {
    {128, 16, 0, 1, 0, 256}, {128, 16, 0, 4, 0, 256}, {128, 32, 0, 1, 0, 256}, {128, 32, 0, 4, 0, 256},
    {128, 64, 0, 4, 0, 256}, {128, 128, 0, 2, 0, 256}, {128, 256, 0, 1, 0, 256},
    {64, 32, 0, 1, 0, 256}, {64, 32, 0, 4, 0, 256}, {64, 8, 0, 1, 0, 256},
    {64, 32, 1, 1, 0, 256}, {64, 32, 1, 4, 0, 256}, {64, 128, 1, 2, 0, 256}, {64, 256, 1, 1, 0, 256},
    {128, 32, 0, 1, 1, 256}, {128, 32, 0, 4, 1, 256},
  };
  const int n = sizeof(h) / sizeof(h[0]);
  Case* d; long long* o; long long ho[64];
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, sizeof(ho)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<<<1, 128, 200 * 1024>>>(d, n, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(ho, o, sizeof(long long) * 2 * n, cudaMemcpyDeviceToHost);
  printf("%5s %5s %6s %5s %3s | %12s %14s\n", "M", "N", "B-maj", "ndst", "TS", "issue cyc/MMA", "done cyc/MMA");
  for (int i = 0; i < n; ++i)
    printf("%5d %5d %6s %5d %3d | %12.1f %14.1f\n", h[i].M, h[i].N, h[i].bmn ? "MN" : "K", h[i].ndst, h[i].ts,
           (double)ho[2 * i] / h[i].reps, (double)ho[2 * i + 1] / h[i].reps);
  return 0;
}
