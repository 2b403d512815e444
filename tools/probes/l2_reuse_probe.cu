// l2_reuse_probe.cu -- how many bytes of a just-streamed tensor can the B200 L2 serve on a second pass?
// (SURVEY 7.1 step 5 asks for this measurement before any "second read of X hits L2" schedule is relied on.)
//
// One cooperative persistent kernel, 2 CTAs per SM.  Phase 1: the grid streams S bytes (ld.global.cg float4, L1 bypassed;
// the grid reads one contiguous moving window: CTA c reads chunk (iter * gridDim + c)).  grid.sync().  Phase 2: the same
// S bytes again, same order, optionally writing S bytes to a second buffer (what stage T does), timed with globaltimer.
// Reported per S: phase-2 read bandwidth.  >> HBM peak = served by L2; ~HBM peak = evicted.
// A "lag" variant re-reads each chunk `lag_bytes` after its first read inside ONE streaming pass (sliding reuse distance),
// which is the access pattern of a pipelined pass-A / stage-T pair.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_reuse_probe l2_reuse_probe.cu && ./l2_reuse_probe
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

constexpr int CHUNK = 64 * 1024;   // bytes per CTA per iteration

__device__ __forceinline__ float read_chunk(const float4* p, int tid, int nthr) {
  float acc = 0.f;
#pragma unroll 4
  for (int i = tid; i < CHUNK / 16; i += nthr) { const float4 v = __ldcg(p + i); acc += v.x + v.y + v.z + v.w; }
  return acc;
}

__global__ void __launch_bounds__(512, 2) two_phase(const float4* x, float4* y, size_t nchunks, int do_write, float* sink, unsigned long long* times) {
  cg::grid_group grid = cg::this_grid();
  float acc = 0.f;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) acc += read_chunk(x + c * (CHUNK / 16), threadIdx.x, blockDim.x);
  grid.sync();
  const unsigned long long t0 = gtime();
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const float4* p = x + c * (CHUNK / 16);
    if (do_write) {
      float4* q = y + c * (CHUNK / 16);
      for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) { float4 v = __ldcg(p + i); v.x *= 2.f; __stcg(q + i, v); }
    } else acc += read_chunk(p, threadIdx.x, blockDim.x);
  }
  grid.sync();
  const unsigned long long t1 = gtime();
  if (blockIdx.x == 0 && threadIdx.x == 0) { times[0] = t0; times[1] = t1; }
  if (acc == 1234.5f) sink[0] = acc;
}

// sliding window: half of the CTAs ("A") stream the tensor, the other half ("T") re-read chunk c once the A side is
// `lag` chunks-of-the-whole-grid ahead (flag per round), optionally writing the same amount elsewhere.
__global__ void __launch_bounds__(512, 1) sliding(const float4* x, float4* y, size_t nchunks, int lag_rounds, int do_write, volatile int* progress,
                                                  float* sink, unsigned long long* times) {
  const int half = gridDim.x / 2;
  const bool is_a = blockIdx.x < half;
  const int me = is_a ? blockIdx.x : blockIdx.x - half;
  float acc = 0.f;
  const unsigned long long t0 = gtime();
  size_t round = 0;
  for (size_t c = me; c < nchunks; c += half, ++round) {
    if (is_a) {
      // flow control: stay at most lag_rounds + 2 ahead of the slowest T CTA (keeps the window bounded)
      if (threadIdx.x == 0) { while ((long long)round - (long long)progress[half + me] > lag_rounds + 2) {} }
      __syncthreads();
      acc += read_chunk(x + c * (CHUNK / 16), threadIdx.x, blockDim.x);
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); progress[me] = (int)round + 1; }
    } else {
      if (threadIdx.x == 0) { while (progress[me] < (int)round + 1 + lag_rounds && progress[me] < (int)((nchunks - me + half - 1) / half)) {} }
      __syncthreads();
      const float4* p = x + c * (CHUNK / 16);
      if (do_write) {
        float4* q = y + c * (CHUNK / 16);
        for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) { float4 v = __ldcg(p + i); v.x *= 2.f; __stcg(q + i, v); }
      } else acc += read_chunk(p, threadIdx.x, blockDim.x);
      __syncthreads();
      if (threadIdx.x == 0) progress[half + me] = (int)round + 1;
    }
  }
  const unsigned long long t1 = gtime();
  if (threadIdx.x == 0) { atomicMin(&times[0], t0); atomicMax(&times[1], t1); }
  if (acc == 1234.5f) sink[0] = acc;
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t maxb = 1024ull << 20;
  float4 *x, *y; float* sink; unsigned long long* times; int* progress;
  cudaMalloc(&x, maxb); cudaMalloc(&y, maxb); cudaMalloc(&sink, 4); cudaMalloc(&times, 16); cudaMalloc(&progress, 4096 * 4);
  cudaMemset(x, 0, maxb); cudaMemset(y, 0, maxb);
  printf("SMs %d\n== two-phase: phase 2 re-reads the S bytes phase 1 streamed (read-only | read + write S)\n", sms);
  printf("%8s %14s %14s\n", "S MB", "reread GB/s", "rd+wr GB/s(rd)");
  for (int mb : {16, 32, 48, 64, 80, 96, 112, 128, 160, 256, 512}) {
    double bw[2];
    for (int w = 0; w < 2; ++w) {
      size_t nchunks = (size_t)mb * (1 << 20) / CHUNK;
      int grid = sms * 2;
      double best = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(y, 0, maxb);                       // flush L2 with unrelated dirty lines, written back before phase 1 ends
        void* args[] = {&x, &y, &nchunks, &w, &sink, &times};
        cudaLaunchCooperativeKernel((void*)two_phase, dim3(grid), dim3(512), args, 0, 0);
        unsigned long long h[2];
        cudaMemcpy(h, times, 16, cudaMemcpyDeviceToHost);
        double gbs = (double)mb * 1.048576e6 / (double)(h[1] - h[0]);
        if (gbs > best) best = gbs;
      }
      bw[w] = best;
    }
    printf("%8d %14.0f %14.0f\n", mb, bw[0], bw[1]);
  }
  printf("== sliding window over a 1 GiB tensor: T CTAs re-read a chunk `lag` MB after the A CTAs streamed it (T also writes: yes)\n");
  printf("%8s %14s %14s\n", "lag MB", "A+T GB/s(X)", "note: X bytes / time; 1 read from HBM + 1 (L2?) re-read + 1 write per byte");
  for (int lag_mb : {0, 4, 8, 16, 24, 32, 48, 64, 96}) {
    int grid = sms / 2 * 2, half = grid / 2;
    size_t nchunks = maxb / CHUNK;
    int lag_rounds = (int)((size_t)lag_mb * (1 << 20) / ((size_t)half * CHUNK));
    int w = 1;
    double best = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(progress, 0, 4096 * 4);
      unsigned long long init[2] = {~0ull, 0ull};
      cudaMemcpy(times, init, 16, cudaMemcpyHostToDevice);
      void* args[] = {&x, &y, &nchunks, &lag_rounds, &w, &progress, &sink, &times};
      cudaLaunchCooperativeKernel((void*)sliding, dim3(grid), dim3(512), args, 0, 0);
      unsigned long long h[2];
      cudaMemcpy(h, times, 16, cudaMemcpyDeviceToHost);
      double gbs = (double)maxb / (double)(h[1] - h[0]);
      if (gbs > best) best = gbs;
    }
    printf("%8d %14.0f   (lag rounds %d; ideal if re-read hits L2: HBM peak / 2 = ~3300; if it misses: / 3 = ~2200)\n", lag_mb, best, lag_rounds);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
