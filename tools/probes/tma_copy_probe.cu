// tma_copy_probe.cu -- ceiling of the stage-T access pattern: persistent CTAs copy X[rows, C] -> Y through shared memory with
// exactly the kernel's tiling (TMA loads of 128 x 32-float slabs into a ring, TMA stores of the same slabs), no compute.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../gansformer-reproducibility-challenge_b200/csrc/gf_tc_common.cuh"
using namespace gf::tc;
namespace gf { void set_error(const char* f, ...) { printf("err %s\n", f); } void set_path(int) {} void note_launch() {} void set_centroid_path(int) {} }

constexpr int SLAB = 16384;
struct Bars { uint64_t full[16], empty[16]; };

__global__ void __launch_bounds__(64, 1) copy_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                                    long long total_tiles, int ns, int nst, int lag) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem);
  const uint32_t ring = smem_u32(smem) + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = (long long)blockIdx.x * total_tiles / gridDim.x, t1 = (long long)(blockIdx.x + 1) * total_tiles / gridDim.x;
  if (threadIdx.x == 0) {                 // producer
    uint32_t ctr = 0;
    for (long long t = t0; t < t1; ++t)
      for (int s = 0; s < ns; ++s, ++ctr) {
        const int st = ctr % nst;
        mbar_wait(smem_u32(&bars->empty[st]), ((ctr / nst) & 1) ^ 1);
        mbar_expect_tx(smem_u32(&bars->full[st]), SLAB);
        tma_load_2d(ring + st * SLAB, &tmX, smem_u32(&bars->full[st]), s * 32, (int)(t * 128));
      }
  } else if (threadIdx.x == 32) {         // storer: keeps `lag` stores in flight before releasing the oldest slab
    uint32_t ctr = 0, rel = 0;
    for (long long t = t0; t < t1; ++t)
      for (int s = 0; s < ns; ++s, ++ctr) {
        const int st = ctr % nst;
        mbar_wait(smem_u32(&bars->full[st]), (ctr / nst) & 1);
        fence_proxy_async();
        tma_store_2d(&tmY, ring + st * SLAB, s * 32, (int)(t * 128));
        tma_commit();
        if (ctr + 1 - rel > (uint32_t)lag) {
          if (lag == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else if (lag == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
          mbar_arrive(smem_u32(&bars->empty[rel % nst])); ++rel;
        }
      }
    tma_wait_all();
  }
}

int main() {
  const int B = 32;
  struct { int n, C; } shapes[] = {{65536, 128}, {16384, 256}, {4096, 512}};
  for (auto sh : shapes) {
    const size_t rows = (size_t)B * sh.n, bytes = rows * sh.C * 4;
    float *x, *y;
    cudaMalloc(&x, bytes * 2); cudaMalloc(&y, bytes);      // two input buffers alternate so reads never hit L2
    cudaMemset(x, 0, bytes * 2);
    CUtensorMap mx[2], my;
    make_map(&mx[0], x, rows, sh.C, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    make_map(&mx[1], x + rows * sh.C, rows, sh.C, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    make_map(&my, y, rows, sh.C, 128, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    const int ns = sh.C / 32;
    for (int nst : {8, 13}) for (int lag : {1, 2, 4}) {
      const int smem = 1024 + nst * SLAB + 1024;
      cudaFuncSetAttribute(copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int iters = 10;
      copy_kernel<<<148, 64, smem>>>(mx[0], my, (long long)rows / 128, ns, nst, lag);
      cudaEventRecord(e0);
      for (int i = 0; i < iters; ++i) copy_kernel<<<148, 64, smem>>>(mx[i & 1], my, (long long)rows / 128, ns, nst, lag);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
      printf("n=%6d C=%3d ring=%2d slabs, %d stores in flight: %.4f ms  %.1f GB/s (read+write)\n", sh.n, sh.C, nst, lag, ms, 2.0 * bytes / ms / 1e6);
    }
    // reference: cudaMemcpyAsync device-to-device of the same bytes
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(y, x + (i & 1) * rows * sh.C, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("n=%6d C=%3d cudaMemcpy D2D: %.4f ms  %.1f GB/s\n", sh.n, sh.C, ms, 2.0 * bytes / ms / 1e6);
    cudaFree(x); cudaFree(y);
  }
  return 0;
}
