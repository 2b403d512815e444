// gf_ops.cu -- memory-bound companions of the attention hot path (include/gf_ops.h): channel scaling
// (style modulation / demodulation), the two upfirdn_2d uses of the generator, and fused bias + noise + activation.
//
// B200-native equivalents of the reference's native ops dnnlib/tflib/ops/{fused_bias_act,upfirdn_2d}.cu (expected
// upstream; not in the checkout).  All are pure streaming kernels: float4 accesses on channels-last rows, grids
// sized to a few waves of 148 SMs, no shared memory (the FIR reuse is served by L1/L2).
#include "gf_common.cuh"
#include "../../include/gf_ops.h"

namespace gf {

static inline int grid_for(size_t work_items, int threads) {
  size_t b = (work_items + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

__global__ void __launch_bounds__(256) chan_scale_kernel(const float4* __restrict__ x, const float4* __restrict__ s,
                                                         float4* __restrict__ y, size_t total4, int hw_c4, int c4n, int s_ld4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw_c4), c4 = (int)(i % c4n);
    const float4 v = x[i], sc = __ldg(s + (size_t)b * s_ld4 + c4);
    y[i] = make_float4(v.x * sc.x, v.y * sc.y, v.z * sc.z, v.w * sc.w);
  }
}

// One thread: a column (w, 4 channels) of ROWS consecutive output rows; horizontal 4-tap pass per input row on the
// fly, vertical pass over a sliding window of 4 horizontally filtered rows.
template <int ROWS>
__global__ void __launch_bounds__(256) blur_up_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                      const float* __restrict__ scale, int Hout, int Wout, int C, float gain) {
  const int c4n = C >> 2;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Wout * c4n) return;
  const int w = t / c4n, c = (t % c4n) * 4;
  const int h0 = blockIdx.y * ROWS, b = blockIdx.z;
  const int Hin = Hout + 1, Win = Wout + 1;
  const float f0 = 0.125f, f1 = 0.375f;
  const float* xb = x + (size_t)b * Hin * Win * C;
  float4 win[4];
  auto hrow = [&](int u) -> float4 {            // u: row of the padded input, x_pad[u][v] = x[u-1][v-1]
    const int r = u - 1;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < 0 || r >= Hin) return a;
    const float* row = xb + (size_t)r * Win * C + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = w + j - 1;
      if (v >= 0 && v < Win) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(row + (size_t)v * C));
        const float f = (j == 0 || j == 3) ? f0 : f1;
        a.x = fmaf(f, q.x, a.x); a.y = fmaf(f, q.y, a.y); a.z = fmaf(f, q.z, a.z); a.w = fmaf(f, q.w, a.w);
      }
    }
    return a;
  };
  win[0] = hrow(h0); win[1] = hrow(h0 + 1); win[2] = hrow(h0 + 2);
  float4 sc = make_float4(gain, gain, gain, gain);
  if (scale) {
    const float4 s = __ldg(reinterpret_cast<const float4*>(scale + (size_t)b * C + c));
    sc = make_float4(gain * s.x, gain * s.y, gain * s.z, gain * s.w);
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int h = h0 + i;
    if (h >= Hout) break;
    win[3] = hrow(h + 3);
    float4 o;
    o.x = (f0 * (win[0].x + win[3].x) + f1 * (win[1].x + win[2].x)) * sc.x;
    o.y = (f0 * (win[0].y + win[3].y) + f1 * (win[1].y + win[2].y)) * sc.y;
    o.z = (f0 * (win[0].z + win[3].z) + f1 * (win[1].z + win[2].z)) * sc.z;
    o.w = (f0 * (win[0].w + win[3].w) + f1 * (win[1].w + win[2].w)) * sc.w;
    *reinterpret_cast<float4*>(y + (((size_t)b * Hout + h) * Wout + w) * C + c) = o;
    win[0] = win[1]; win[1] = win[2]; win[2] = win[3];
  }
}

__global__ void __launch_bounds__(256) upsample2x_nchw_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                              float* __restrict__ y, int planes, int H, int W) {
  const int OW = 2 * W, OH = 2 * H;
  const size_t total = (size_t)planes * OH * OW;
  const float f[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const size_t pl = i / ((size_t)OW * OH);
    const float* xp = x + pl * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int u = oy + a - 2;                   // row in the zero-inserted image
      if (u < 0 || (u & 1) || (u >> 1) >= H) continue;
#pragma unroll
      for (int bq = 0; bq < 4; ++bq) {
        const int v = ox + bq - 2;
        if (v < 0 || (v & 1) || (v >> 1) >= W) continue;
        acc = fmaf(f[a] * f[bq], xp[(size_t)(u >> 1) * W + (v >> 1)], acc);
      }
    }
    acc *= 4.f;
    y[i] = add ? acc + add[i] : acc;
  }
}

__global__ void __launch_bounds__(256) bias_act_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                       const float* __restrict__ bias, const float* __restrict__ noise,
                                                       const float* __restrict__ strength, long long noise_bstride,
                                                       size_t total4, int HW, int c4n, int act, float gain) {
  const float st = (noise && strength) ? __ldg(strength) : (noise ? 1.f : 0.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const size_t tok = i / c4n;
    const int t = (int)(tok % HW);
    const size_t b = tok / HW;
    float4 v = x[i];
    float nz = 0.f;
    if (noise) nz = __ldg(noise + b * noise_bstride + t) * st;
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bb = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    v.x += nz + bb.x; v.y += nz + bb.y; v.z += nz + bb.z; v.w += nz + bb.w;
    if (act == 1) {
      v.x = fmaxf(v.x, 0.2f * v.x); v.y = fmaxf(v.y, 0.2f * v.y); v.z = fmaxf(v.z, 0.2f * v.z); v.w = fmaxf(v.w, 0.2f * v.w);
    }
    y[i] = make_float4(v.x * gain, v.y * gain, v.z * gain, v.w * gain);
  }
}

// d[b,o] = rsqrt(sum_i s[b,i]^2 * wsq[o,i] + eps): one warp per output, lanes stride over i
__global__ void __launch_bounds__(256) demod_coef_kernel(const float* __restrict__ s, const float* __restrict__ wsq,
                                                         float* __restrict__ d, int B, int O, int I, float eps, int s_ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * O) return;
  const int b = warp / O, o = warp % O;
  const float* sb = s + (size_t)b * s_ld;
  const float* wo = wsq + (size_t)o * I;
  float acc = 0.f;
  for (int i = lane; i < I; i += 32) { const float v = sb[i]; acc = fmaf(v * v, wo[i], acc); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) d[warp] = rsqrtf(acc + eps);
}

}  // namespace gf

using namespace gf;

extern "C" {

int gf_chan_scale_nhwc(const float* x, const float* s, int s_ld, float* y, int B, int HW, int C, void* stream) {
  if (!x || !s || !y) { set_error("gf_chan_scale_nhwc: null pointer"); return GF_ERR_INVALID; }
  if (B <= 0 || HW <= 0 || C <= 0 || (C & 3) || s_ld < C || (s_ld & 3) || ((uintptr_t)s & 15)) {
    set_error("gf_chan_scale_nhwc: need B,HW,C > 0, C %% 4 == 0, s_ld >= C, s_ld %% 4 == 0, s 16-byte aligned (C=%d s_ld=%d)", C, s_ld);
    return GF_ERR_UNSUPPORTED;
  }
  const size_t total4 = (size_t)B * HW * (C >> 2);
  chan_scale_kernel<<<grid_for(total4, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(s), reinterpret_cast<float4*>(y), total4, HW * (C >> 2), C >> 2, s_ld >> 2);
  GF_LAUNCH_OK();
  return GF_OK;
}

int gf_blur_up_nhwc(const float* x, float* y, const float* scale, int B, int Hout, int Wout, int C, float gain, void* stream) {
  if (!x || !y) { set_error("gf_blur_up_nhwc: null pointer"); return GF_ERR_INVALID; }
  if (B <= 0 || Hout <= 0 || Wout <= 0 || C <= 0 || (C & 3) || B > 65535) { set_error("gf_blur_up_nhwc: bad shape (B=%d Hout=%d Wout=%d C=%d)", B, Hout, Wout, C); return GF_ERR_UNSUPPORTED; }
  constexpr int ROWS = 8;
  dim3 grid((Wout * (C >> 2) + 255) / 256, (Hout + ROWS - 1) / ROWS, B);
  blur_up_kernel<ROWS><<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, scale, Hout, Wout, C, gain);
  GF_LAUNCH_OK();
  return GF_OK;
}

int gf_upsample2x_nchw(const float* x, const float* add, float* y, int B, int C, int H, int W, void* stream) {
  if (!x || !y) { set_error("gf_upsample2x_nchw: null pointer"); return GF_ERR_INVALID; }
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) { set_error("gf_upsample2x_nchw: bad shape"); return GF_ERR_INVALID; }
  const size_t total = (size_t)B * C * 4 * H * W;
  upsample2x_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, add, y, B * C, H, W);
  GF_LAUNCH_OK();
  return GF_OK;
}

int gf_bias_act_nhwc(const float* x, float* y, const float* bias, const float* noise, const float* strength,
                     long long noise_bstride, int B, int HW, int C, int act, float gain, void* stream) {
  if (!x || !y) { set_error("gf_bias_act_nhwc: null pointer"); return GF_ERR_INVALID; }
  if (B <= 0 || HW <= 0 || C <= 0 || (C & 3) || act < 0 || act > 1) { set_error("gf_bias_act_nhwc: bad arguments (C=%d act=%d)", C, act); return GF_ERR_UNSUPPORTED; }
  const size_t total4 = (size_t)B * HW * (C >> 2);
  bias_act_kernel<<<grid_for(total4, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), bias, noise, strength, noise_bstride, total4, HW, C >> 2, act, gain);
  GF_LAUNCH_OK();
  return GF_OK;
}

int gf_demod_coef(const float* styles, int s_ld, const float* wsq, float* d, int B, int O, int I, float eps, void* stream) {
  if (!styles || !wsq || !d) { set_error("gf_demod_coef: null pointer"); return GF_ERR_INVALID; }
  if (B <= 0 || O <= 0 || I <= 0 || s_ld < I) { set_error("gf_demod_coef: bad shape"); return GF_ERR_INVALID; }
  const long long warps = (long long)B * O;
  demod_coef_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(styles, wsq, d, B, O, I, eps, s_ld);
  GF_LAUNCH_OK();
  return GF_OK;
}

}  // extern "C"
