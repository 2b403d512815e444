// gf_common.cuh -- shared host/device definitions for libgf_attn (sm_100a only).
//
// Buffer layouts are the ones restated in oracle/folded.py (stage W / I / T); keep the two in sync.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/gf_attn.h"

namespace gf {

// ---- thread-local error reporting -----------------------------------------------------------------
void set_error(const char* fmt, ...);
void set_path(int path);
void set_centroid_path(int path);
void note_launch();

#define GF_CUDA_OK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      gf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return GF_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

#define GF_LAUNCH_OK()                                                                           \
  do {                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) {                                                                     \
      gf::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));   \
      return GF_ERR_CUDA;                                                                        \
    }                                                                                            \
    gf::note_launch();                                                                           \
  } while (0)

// ---- layout of the folded-weight buffer and of the per-call workspace -----------------------------
// All offsets in floats, each region 64-float (256 B) aligned.
struct Layout {
  int B, H, W, C, k, D, p, KP, Cout, LDK, n, duplex;
  // folded buffer (stage W)
  size_t f_AK, f_CK, f_AV, f_CV, f_ROW, f_COL;        // simplex + duplex
  size_t f_WV2, f_BV2, f_AM, f_CM;                    // duplex only
  size_t f_QFOLD, f_KCONST, f_MFOLD, f_QCONST;        // scratch of stage W
  size_t f_total;
  // workspace (stage I/T)
  size_t w_KPALL, w_Kp, w_Vt, w_Rt, w_Ct;             // keys / values / positional logit tables
  size_t w_NSCALE, w_NSHIFT, w_NPART;                 // instance/batch norm statistics
  size_t w_MALL, w_M, w_Rt2, w_Ct2, w_PART, w_XBAR;   // duplex pass A
  size_t f_AK2, f_CK2;                                // duplex: keys straight from Xbar (Wv2 and bv2 folded into AK / CK)
  size_t f_ACQ, f_WI2L, f_BI2L;                       // kmeans_iters > 1: centroid -> pass-A query table; g_img2ltnt: centroid -> latent gain
  size_t w_CEN, w_Y2;                                 // scratch centroids [B,k,C] (caller passed none), modulated latents [B,k,D]
  int iters, img2ltnt;
  int heads, seg;                                     // num_heads; per-head segment of the KP "latent" columns (KP = heads * seg, seg >= k)
  size_t w_total;
  int nsplit_norm, nsplit_cen;
};

// SM count of the current device (B200: 148), queried once per process; grids and split models are sized from it.
inline int num_sms() {
  static int v = -1;
  if (v < 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) v = n;
    else return 148;          // no device visible (host-only layout queries): the B200 figure, not cached
  }
  return v;
}

inline size_t align64(size_t x) { return (x + 63) & ~size_t(63); }
inline int pad_k(int k) { return k <= 16 ? 16 : 32; }

// Fills L; returns GF_OK or an error (message set).
int make_layout(const gf_attn_desc* d, Layout* L);
int check_device();          // GF_OK on a compute-capability-10.x device, an error otherwise (gf_api.cu)

// ---- stage W / I kernels (gf_fold.cu) ---------------------------------------------------------------
int fold_weights(const Layout& L, const gf_attn_desc* d, const gf_attn_weights* w, float* folded, cudaStream_t st);
// key_source: Y [B*k, D] (simplex) or centroids [B*k, C] (duplex); kdim = D or C.
int prologue(const Layout& L, const gf_attn_desc* d, const float* Y, const float* key_source, int kdim,
             const float* folded, float* ws, cudaStream_t st, const float* in_scale = nullptr, int in_scale_ld = 0,
             bool keys_from_xbar = false, bool with_v = true);
int prologue_batch(int n, const Layout* Ls, const gf_attn_desc* const* ds, const float* const* Ys, const float* const* fs, float* const* wss,
                   const gf_attn_postop* const* posts, cudaStream_t st);
int duplex_tables_from_centroids(const Layout& L, const gf_attn_desc* d, const float* cen, const float* Y, const float* f, float* ws,
                                 cudaStream_t st, const float* in_scale, int in_scale_ld);
int img2ltnt(const Layout& L, const float* Y, const float* cen, const float* f, float* ws, cudaStream_t st);
int duplex_tables(const Layout& L, const gf_attn_desc* d, const float* Y, const float* folded, float* ws, cudaStream_t st,
                  const float* in_scale = nullptr, int in_scale_ld = 0);
// C[M,N] = alpha * opA(A) opB(B) + E[(m % emod), n] + v[n]
int gemm(cudaStream_t st, int M, int N, int K, const float* A, int lda, bool ta, const float* B, int ldb, bool tb,
         float* Cm, int ldc, float alpha, const float* E = nullptr, int lde = 0, int emod = 1, const float* v = nullptr,
         bool allow_tf32 = false);
// tcgen05 TF32 version for dense row-major operands (gf_tc_gemm.cu); gemm() routes to it when allow_tf32 and the shape fits
bool gemm_tc_ok(int M, int N, int K, const float* A, const float* B, const float* Cm, int ldc);
int gemm_tc(cudaStream_t st, int M, int N, int K, const float* A, const float* B, float* Cm, int ldc, float alpha,
            const float* E, int lde, int emod, const float* v);

// ---- stage T kernels ----------------------------------------------------------------------------------
int token_pass_simt(const Layout& L, const gf_attn_desc* d, const float* X, float* Xout, float* att, float* ws, const gf_attn_postop* post, cudaStream_t st);
int norm_stats(const Layout& L, const gf_attn_desc* d, const float* X, float* ws, cudaStream_t st);
int centroid_pass_simt(const Layout& L, const gf_attn_desc* d, const float* X, float* ws, cudaStream_t st,
                       const float* in_scale = nullptr, int in_scale_ld = 0);
// Xbar = merge of the split partials (times the load-side scale, when given)
int centroid_merge(const Layout& L, float* ws, cudaStream_t st, const float* in_scale = nullptr, int in_scale_ld = 0);
// tcgen05 duplex pass A (gf_tc_cen.cu): partials into ws (same format as the CUDA-core kernel), then centroid_merge
bool tc_centroid_supported(const Layout& L, const gf_attn_desc* d);
// in_scale: only used when the split count is 1 and the kernel writes the normalised Xbar itself (no merge kernel)
int centroid_pass_tc(const Layout& L, const gf_attn_desc* d, const float* X, float* ws, cudaStream_t st,
                     const float* in_scale = nullptr, int in_scale_ld = 0);
// tcgen05 / TMA path (gf_tc.cu).  tc_supported() says whether the shape is served by it.
bool tc_supported(const Layout& L, const gf_attn_desc* d);
int token_pass_tc(const Layout& L, const gf_attn_desc* d, const float* X, float* Xout, float* att, float* ws, const gf_attn_postop* post, cudaStream_t st);

}  // namespace gf
