// gf_tc_cen.cu -- duplex pass A (latents attend to the grid, softmax over the n grid cells) on the tensor path.
//
// Replaces, on the reference side (expected src/training/network.py, not in the checkout): the k-means / centroid branch
// of transformer_layer (image -> latents attention).  Algorithm = oracle/folded.py centroid_pass():
//     L[t,j] = x_t . M_j + pos(t,j);   A[j,:] = softmax_t L[:,j];   Xbar[j,:] = sum_t A[j,t] x_t
// streamed once over X with an online softmax (lazy rescaling), split over `nsplit` CTAs per image whose partials
// (acc[KP][C], m[KP], l[KP]) are merged by centroid_merge_kernel (gf_simt.cu).
//
// grid (nsplit, B), 6 warps:
//   warp 0     TMA producer: M (latent-query matrix, once) and the X slabs (128 tokens x 32 channels).  Every slab is
//              fetched twice: with SWIZZLE_128B for GEMM1 (X is the K-major A operand: contraction over channels) and with
//              SWIZZLE_128B_ATOM_32B for GEMM2 (X is the MN-major B operand: contraction over tokens; for 32-bit MN-major
//              operands that swizzle -- UMMA layout SWIZZLE_128B_BASE32B -- is the only one the tensor core accepts).
//              The second fetch hits L2; slot order in the ring: P1(0), [P1(i+1), P2(i)] for i = 0, 1, ...
//   MMA issuers: one thread retires ~10 scalar instructions (~90 cycles) per tcgen05.mma whatever its shape
//              (tools/probes/mma_probe.cu), and a tile needs 4*C/32 + 16*C/32 + 16 of them, so the issue is spread over warps:
//   warp 1     GEMM1  S[128 tok, KP]   = X . M^T                      (M=128, as in stage T), runs one tile ahead
//   warps 6-9  GEMM2  D2[KP->64, 32ch] += E^T[64, 128 tok] . X_slab   (M=64; A = E^T from smem, K-major); slab s -> warp 6 + s%4
//   warp 10    GEMM3  D3[64, 8]        += E^T . 1                      (softmax denominators)
//              Every issuer walks ALL ring fills in slot order and arrives on slab_empty for each (the owner through
//              tcgen05.commit, the others with a plain arrive; count = number of issuers): a stage is only refilled once every
//              issuer has seen its current fill, so no parity wait can fall two phases behind its barrier.
//   warps 2-5  row warps (thread = token): positional logits, S from TMEM, per-latent tile maximum (warp shuffles +
//              shared memory), E = exp(S - m) rounded to TF32 and written TRANSPOSED into shared memory, lazy rescale of
//              the TMEM accumulators when a running maximum moves by more than TAU, final flush of the partials.
// HBM traffic: X read once (+ one L2 re-read).
#include <stdlib.h>
#include "gf_common.cuh"
#include "gf_tc_common.cuh"

namespace gf {
namespace tcc {

using namespace tc;

constexpr int TILE = 128;
constexpr int SLAB_CH = 32;
constexpr int SLAB_BYTES = TILE * SLAB_CH * 4;
constexpr int MAX_STAGES = 12;
constexpr int NUM_THREADS = 352;     // warp 0 producer, 1 GEMM1, 2-5 row warps, 6-9 GEMM2 issuers, 10 GEMM3 (denominators)
constexpr int TMEM_COLS = 512;
constexpr int COL_S = 0;          // S[2]: 2 x 32 columns
constexpr int COL_D3 = 64;        // softmax denominators (8 columns used)
constexpr int COL_D2 = 128;       // Xbar accumulators: C columns (C <= 256)
constexpr float TAU = 8.f;        // lazy rescale threshold (natural-log units): exp(8) ~ 3e3 of headroom is harmless in fp32

struct Params {
  const float* Rt; const float* Ct; float* part;
  int n, H, W, k, nsplit, tiles_per_image, nstages;
};

struct Bars {
  uint64_t slab_full[MAX_STAGES], slab_empty[MAX_STAGES];
  uint64_t m_full, done;
  uint64_t s_full[2], e_full[2], e_free[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// NS = slabs of GEMM1 (all C channels); NS2 = slabs of GEMM2 handled by this CTA (C/32 or, for C = 512, half of them:
// blockIdx.z selects the channel half, both CTAs recompute the cheap GEMM1 + softmax; TMEM holds NS2*32 accumulator columns)
template <int KP, int NS, int NS2 = (NS > 8 ? NS / 2 : NS)>
struct Cfg {
  static constexpr int C = NS * SLAB_CH;
  static constexpr int M_BYTES = KP * C * 4;                 // NS chunks of [KP rows x 128 B]
  static constexpr int E_CHUNK = KP * 128;                   // one 32-token chunk of E^T: [KP rows x 128 B]
  static constexpr int E_BYTES = 4 * E_CHUNK;                // 128 tokens
  static constexpr int OFF_M = 0;
  static constexpr int OFF_E = OFF_M + M_BYTES;              // 2 buffers
  static constexpr int OFF_ONES = OFF_E + 2 * E_BYTES;       // 1 KB of 1.0f
  static constexpr int OFF_SMALL = OFF_ONES + 1024;          // red[4][KP], mref[KP], resc[KP], flag
  static constexpr int SMALL_BYTES = (4 * KP + 2 * KP + 4) * 4;
  static constexpr int OFF_BARS = (OFF_SMALL + SMALL_BYTES + 15) / 16 * 16;
  // the M=64 A descriptor of GEMM2 reads 64 rows from each E chunk: rows >= KP alias whatever follows (their D rows are
  // never read); the ring behind keeps those reads inside the allocation.
  static constexpr int OFF_RING = (OFF_BARS + (int)sizeof(Bars) + 1023) / 1024 * 1024;
  static constexpr int FIXED_BYTES = OFF_RING;
};

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major 32-bit operand, UMMA layout SWIZZLE_128B_BASE32B (= TMA SWIZZLE_128B_ATOM_32B, cute Swizzle<2,5,2>):
// 32 contiguous MN elements (128 B) per K row, 4 K rows per 512-byte atom (32-byte chunks XOR row % 4);
// leading byte offset = next block of 32 MN elements, stride byte offset = next 4 K rows.
constexpr uint32_t LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)LAYOUT_SW128_BASE32B << 61;
  return d;
}

__device__ __forceinline__ int ntiles_dbg(const Params& P, int sp) { const int per = (P.tiles_per_image + P.nsplit - 1) / P.nsplit; return min(P.tiles_per_image, sp * per + per) - sp * per; }

template <int KP, int NS, int NS2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
centroid_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                   const __grid_constant__ CUtensorMap tmM, const Params P) {
  using CF = Cfg<KP, NS, NS2>;
  constexpr int C = CF::C;
  constexpr int C2 = NS2 * SLAB_CH;                  // accumulator channels of this CTA
  constexpr int NG2 = NS2 < 4 ? NS2 : 4;             // GEMM2 issuer warps
  const int zoff = blockIdx.z * C2;                   // first channel of this CTA's GEMM2 share
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_m = s_base + CF::OFF_M, s_e = s_base + CF::OFF_E, s_ones = s_base + CF::OFF_ONES, s_ring = s_base + CF::OFF_RING;
  Bars* bars = reinterpret_cast<Bars*>(smem + CF::OFF_BARS);
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && g_dbg_buf) { g_dbg_buf[1] = smem_u32(bars); g_dbg_buf[2] = (unsigned)ntiles_dbg(P, blockIdx.x); }
  float* red = reinterpret_cast<float*>(smem + CF::OFF_SMALL);      // [4][KP]
  float* mref = red + 4 * KP;                                       // [KP]
  float* resc = mref + KP;                                          // [KP]
  volatile int* trigf = reinterpret_cast<volatile int*>(resc + KP);  // [2] per-tile-parity trigger flags
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nst = P.nstages;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int per = (P.tiles_per_image + P.nsplit - 1) / P.nsplit;
  const int tile_beg = sp * per, tile_end = min(P.tiles_per_image, tile_beg + per);
  const int ntiles = tile_end - tile_beg;
  float* part = P.part + ((size_t)b * P.nsplit + sp) * KP * (C + 4);

  if (ntiles <= 0) {                          // empty split: neutral partial
    for (int i = threadIdx.x; i < KP * (C2 + 4); i += NUM_THREADS) {
      const int j = i / (C2 + 4), c = i % (C2 + 4);
      if (c < C2) part[(size_t)j * (C + 4) + zoff + c] = 0.f;
      else if (blockIdx.z == 0) part[(size_t)j * (C + 4) + C + (c - C2)] = (c == C2) ? -INFINITY : 0.f;
    }
    return;
  }

  for (int i = threadIdx.x; i < 256; i += NUM_THREADS) reinterpret_cast<float*>(smem + CF::OFF_ONES)[i] = 1.f;
  if (threadIdx.x < KP) { mref[threadIdx.x] = -INFINITY; resc[threadIdx.x] = 1.f; }
  if (threadIdx.x < 2) trigf[threadIdx.x] = 0;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmX2); prefetch_tmap(&tmM);
    for (int i = 0; i < nst; ++i) { mbar_init(smem_u32(&bars->slab_full[i]), 1); mbar_init(smem_u32(&bars->slab_empty[i]), NG2 + 2); }
    mbar_init(smem_u32(&bars->m_full), 1); mbar_init(smem_u32(&bars->done), NG2 + 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->e_full[i]), 4);
      mbar_init(smem_u32(&bars->e_free[i]), NG2 + 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();                        // the generic-proxy writes of `ones` must be visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const uint32_t mb = smem_u32(&bars->m_full);
      mbar_expect_tx(mb, (uint32_t)CF::M_BYTES);
#pragma unroll
      for (int s = 0; s < NS; ++s) tma_load_2d(s_m + s * (KP * 128), &tmM, mb, s * SLAB_CH, b * KP);
      uint32_t ctr = 0;
      auto load_tile = [&](int it, const CUtensorMap* map, int nslabs, int c_first) {
        const int row0 = (b * P.tiles_per_image + tile_beg + it) * TILE;
        for (int s = 0; s < nslabs; ++s, ++ctr) {
          const int stage = (int)(ctr % (uint32_t)nst);
          mbar_wait(smem_u32(&bars->slab_empty[stage]), ((ctr / (uint32_t)nst) & 1u) ^ 1u);
          const uint32_t bar = smem_u32(&bars->slab_full[stage]);
          mbar_expect_tx(bar, SLAB_BYTES);
          tma_load_2d(s_ring + stage * SLAB_BYTES, map, bar, c_first + s * SLAB_CH, row0);
        }
      };
      load_tile(0, &tmX, NS, 0);
      for (int it = 0; it < ntiles; ++it) {
        if (it + 1 < ntiles) load_tile(it + 1, &tmX, NS, 0);   // GEMM1 of the next tile runs ahead of GEMM2 of this one
        load_tile(it, &tmX2, NS2, zoff);                        // second fetch (L2), MN-major swizzle, this CTA's channels
      }
    }
  } else if (warp == 1 || warp >= 6) {
    // =============================== MMA issuers ===============================
    // ring slot order: P1(0) | P1(1) P2(0) | P1(2) P2(1) | ... ; slots before block `it`: NS + it*(NS+NS2)
    const int role = warp == 1 ? 0 : (warp == 10 ? 2 : 1);         // 0: GEMM1, 1: GEMM2 share (warp - 6), 2: GEMM3
    const int g2r = warp - 6;
    if (lane == 0 && !(role == 1 && g2r >= NG2)) {
      constexpr uint32_t IDESC1 = idesc_tf32(TILE, KP, 0);
      constexpr uint32_t IDESC2 = idesc_tf32(64, 32, 1);
      constexpr uint32_t IDESC3 = idesc_tf32(64, 8, 0);
      mbar_wait(smem_u32(&bars->m_full), 0);
      tc_fence_after();
      // Descriptors are built once and advanced with one 64-bit add per MMA (start-address field = 16-byte units).
      const uint64_t dM0 = umma_desc(s_m, 1024, LAYOUT_SW128);
      const uint64_t dOnes = umma_desc(s_ones, 1024, LAYOUT_SW128);
      const uint64_t dRingK = umma_desc(s_ring, 1024, LAYOUT_SW128);             // slab as K-major A (GEMM1)
      const uint64_t dRingMN = umma_desc_mn(s_ring, SLAB_BYTES, 512);            // slab as MN-major B (GEMM2)
      const uint64_t dE0 = umma_desc(s_e, 1024, LAYOUT_SW128);
      uint32_t ctr = 0;
      auto wait_slot = [&]() -> int {                                             // observe the next ring fill
        const int stage = (int)(ctr % (uint32_t)nst);
        mbar_wait(smem_u32(&bars->slab_full[stage]), (ctr / (uint32_t)nst) & 1u);
        ++ctr;
        return stage;
      };
      auto pass1 = [&](int it) {                                                  // the NS fills of P1(it)
        const uint32_t d_s = tmem + COL_S + (it & 1) * 32;
        for (int s = 0; s < NS; ++s) {
          const int stage = wait_slot();
          if (role == 0) {
            tc_fence_after();
            const uint64_t da = dRingK + (uint64_t)(stage * (SLAB_BYTES >> 4));
            const uint64_t db = dM0 + (uint64_t)(s * ((KP * 128) >> 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(d_s, da + kk * 2, db + kk * 2, IDESC1, (s | kk) ? 1u : 0u);
            umma_commit(smem_u32(&bars->slab_empty[stage]));
          } else {
            mbar_arrive(smem_u32(&bars->slab_empty[stage]));         // seen
          }
        }
        if (role == 0) umma_commit(smem_u32(&bars->s_full[it & 1]));
      };
      pass1(0);
      for (int it = 0; it < ntiles; ++it) {
        const int buf = it & 1;
        if (role == 0 && it >= 1) {
          // S[(it+1)&1] was read by the row warps for tile it-1; they arrive on e_full(it-1) after reading it.  (Waiting it
          // every iteration also keeps this thread's parity bookkeeping of e_full in step.)
          mbar_wait(smem_u32(&bars->e_full[buf ^ 1]), (uint32_t)(((it - 1) >> 1) & 1));
          tc_fence_after();
        }
        if (it + 1 < ntiles) pass1(it + 1);                          // GEMM1 of the next tile keeps the row warps fed
        const uint64_t de = dE0 + (uint64_t)(buf * (CF::E_BYTES >> 4));
        const uint32_t acc0 = it ? 1u : 0u;
        if (role != 0) {
          mbar_wait(smem_u32(&bars->e_full[buf]), (uint32_t)((it >> 1) & 1));
          tc_fence_after();
        }
        if (role == 2) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)
            umma_ss(tmem + COL_D3, de + (uint64_t)(((kk >> 2) * CF::E_CHUNK + (kk & 3) * 32) >> 4), dOnes, IDESC3, kk ? 1u : acc0);
        }
        for (int s = 0; s < NS2; ++s) {                              // the NS2 fills of P2(it)
          const int stage = wait_slot();
          if (role == 1 && (s % NG2) == g2r) {
            tc_fence_after();
            const uint64_t dx = dRingMN + (uint64_t)(stage * (SLAB_BYTES >> 4));
            const uint32_t d2 = tmem + COL_D2 + s * 32;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)                         // 8 tokens (two 4-row swizzle atoms) per MMA
              umma_ss(d2, de + (uint64_t)(((kk >> 2) * CF::E_CHUNK + (kk & 3) * 32) >> 4), dx + (uint64_t)(kk * 64), IDESC2,
                      kk ? 1u : acc0);
            umma_commit(smem_u32(&bars->slab_empty[stage]));        // slab recycled once this issuer's MMAs are done
          } else {
            mbar_arrive(smem_u32(&bars->slab_empty[stage]));         // seen
          }
        }
        if (role != 0) umma_commit(smem_u32(&bars->e_free[buf]));
      }
      if (role != 0) umma_commit(smem_u32(&bars->done));
    }
  } else {
    // =============================== row warps ===============================
    // One warp per scheduler and nothing to hide latency with: the loop is written for few instructions and no exposed
    // loads -- next tile's positional logits are prefetched, references live in registers (pre-multiplied by log2 e),
    // exp is one ex2.approx, the swizzled store offsets are precomputed.
    const int q = warp & 3;                                         // TMEM lane quadrant == 32-token chunk of the tile
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int rtid = (warp - 2) * 32 + lane;                        // 0..127 inside the row-warp group
    constexpr float LOG2E = 1.4426950408889634f;
    auto load_pos = [&](int it, float* dst) {
      const int tok = (tile_beg + it) * TILE + row;
      const int h = tok / P.W, w = tok - h * P.W;
      const float4* rt = reinterpret_cast<const float4*>(P.Rt + ((size_t)b * P.H + h) * KP);
      const float4* ct = reinterpret_cast<const float4*>(P.Ct + ((size_t)b * P.W + w) * KP);
#pragma unroll
      for (int j4 = 0; j4 < KP / 4; ++j4) {
        const float4 r = __ldg(rt + j4), c = __ldg(ct + j4);
        dst[j4 * 4 + 0] = r.x + c.x; dst[j4 * 4 + 1] = r.y + c.y; dst[j4 * 4 + 2] = r.z + c.z; dst[j4 * 4 + 3] = r.w + c.w;
      }
    };
    float pos[KP];                                                  // positional logits of the current tile
    float ml[KP];                                                   // running references * log2(e); padded latents: 0
    load_pos(0, pos);
#pragma unroll
    for (int j = 0; j < KP; ++j) ml[j] = j < P.k ? -INFINITY : 0.f;
    uint32_t soff[8];                                               // byte offset of this token inside row j of a chunk
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) soff[jj] = (uint32_t)((((lane >> 2) ^ jj) << 4) + (lane & 3) * 4);
    for (int it = 0; it < ntiles; ++it) {
      const int buf = it & 1;
      const uint32_t bph = (uint32_t)((it >> 1) & 1);
      mbar_wait(smem_u32(&bars->s_full[buf]), bph);
      tc_fence_after();
      float sv[KP];
      tmem_ld16(tmem + lane_addr + COL_S + buf * 32, sv);
      if constexpr (KP == 32) tmem_ld16(tmem + lane_addr + COL_S + buf * 32 + 16, sv + 16);
      tmem_wait_ld();
      // t_j = (logit - reference) * log2(e); the trigger test needs only its maximum over this thread's latents
      float ex = -INFINITY;
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        sv[j] = (sv[j] + pos[j]) * LOG2E;                           // logits in log2 units (padded latents: -inf)
        ex = fmaxf(ex, sv[j] - ml[j]);                              // first tile: +inf -> trigger
      }
      if (it + 1 < ntiles) load_pos(it + 1, pos);                   // prefetch: consumed one tile later
      const bool trig = __any_sync(0xffffffffu, ex > TAU * LOG2E);
      if (lane == 0 && trig) trigf[buf] = 1;
      if (rtid == 0) trigf[buf ^ 1] = 0;                            // clean flag for the next tile
      named_bar_sync(1, 128);
      const bool full = trigf[buf] != 0;
      if (full) {   // rare: some logit jumped more than TAU above its reference (always on the first tile)
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          float v = sv[j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
          if (lane == 0) red[(warp - 2) * KP + j] = v;
        }
        named_bar_sync(1, 128);
        if (rtid < KP) {
          const float tm = fmaxf(fmaxf(red[rtid], red[KP + rtid]), fmaxf(red[2 * KP + rtid], red[3 * KP + rtid]));
          const float mo = mref[rtid];                              // log2 units; -inf before the first tile
          float mn = mo, f = 1.f;
          if (rtid < P.k && (tm > mo + TAU * LOG2E || mo == -INFINITY)) {
            mn = tm;
            f = (mo == -INFINITY) ? 1.f : exp2f(mo - mn);
          }
          mref[rtid] = mn;
          resc[rtid] = f;
        }
        named_bar_sync(1, 128);
#pragma unroll
        for (int j = 0; j < KP; ++j) ml[j] = j < P.k ? mref[j] : 0.f;
      }
      // ---- E = 2^(t_j), written transposed (E^T[latent][token], K-major SW128, 32-token chunks).  E is NOT rounded to
      //      TF32: the tensor core's truncation bias hits numerator (D2) and denominator (D3) alike and cancels.
      // buffer `buf` was last read by GEMM2(it-2), whose completion this thread observed during tile it-1 (below)
      uint8_t* eb = smem + CF::OFF_E + buf * CF::E_BYTES + q * CF::E_CHUNK;
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(sv[j] - ml[j]));
        *reinterpret_cast<float*>(eb + j * 128 + soff[j & 7]) = e;
      }
      // ---- every tile: observe the completion of GEMM2(it-1) (parity bookkeeping must not skip phases); then, if a
      //      running maximum moved, rescale the accumulators before GEMM2(it) adds to them
      if (it > 0) {
        mbar_wait(smem_u32(&bars->e_free[buf ^ 1]), (uint32_t)(((it - 1) >> 1) & 1));
        tc_fence_after();
      }
      if (full && it > 0) {
        // M=64 accumulators: latent j lives in TMEM lane (j % 16) + 32 * (j / 16): lanes 0-15 of quadrants 0 (and 1)
        if (q * 16 < KP) {
          const float f = lane < 16 ? resc[q * 16 + lane] : 1.f;
          float v[16];
#pragma unroll 1
          for (int c0 = 0; c0 < C2; c0 += 16) {
            tmem_ld16(tmem + lane_addr + COL_D2 + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= f;
            tmem_st16(tmem + lane_addr + COL_D2 + c0, v);
          }
          tmem_ld16(tmem + lane_addr + COL_D3, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= f;
          tmem_st16(tmem + lane_addr + COL_D3, v);
          tmem_wait_st();
        }
      }
      fence_proxy_async();                                          // E^T (generic proxy) -> tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->e_full[buf]));
    }
    // ---- flush: partial accumulators, running maxima and denominators of this split
    mbar_wait(smem_u32(&bars->done), 0);
    tc_fence_after();
    if (q * 16 < KP) {
      const int j = q * 16 + lane;                                  // valid for lane < 16
      float v[16];
#pragma unroll 1
      for (int c0 = 0; c0 < C2; c0 += 16) {
        tmem_ld16(tmem + lane_addr + COL_D2 + c0, v);
        tmem_wait_ld();
        if (lane < 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) part[(size_t)j * (C + 4) + zoff + c0 + i] = v[i] * 1.000352220f;   // X truncation bias (gf_fold.cu)
        }
      }
      tmem_ld16(tmem + lane_addr + COL_D3, v);
      tmem_wait_ld();
      if (lane < 16 && blockIdx.z == 0) {
        part[(size_t)j * (C + 4) + C] = j < P.k ? mref[j] * 0.6931471805599453f : -INFINITY;   // log2 -> natural units
        part[(size_t)j * (C + 4) + C + 1] = v[0];
        part[(size_t)j * (C + 4) + C + 2] = 0.f;
        part[(size_t)j * (C + 4) + C + 3] = 0.f;
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

template <int KP, int NS>
static int stages_for(int smem_limit) {
  int st = (smem_limit - Cfg<KP, NS>::FIXED_BYTES - 1024) / SLAB_BYTES;
  return st > MAX_STAGES ? MAX_STAGES : st;
}

template <int KP, int NS>
static int launch(const Layout& L, const float* X, float* ws, cudaStream_t st) {
  using CF = Cfg<KP, NS>;
  constexpr int NS2 = NS > 8 ? NS / 2 : NS;
  const int nst = stages_for<KP, NS>(device_smem_optin());
  if (nst < 4) { set_error("tcgen05 centroid pass: shared memory too small for C=%d KP=%d", L.C, KP); return GF_ERR_UNSUPPORTED; }
  CUtensorMap tmX, tmX2, tmM;
  int rc;
  if ((rc = make_map(&tmX, X, (uint64_t)L.B * L.n, L.C, TILE, SLAB_CH, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&tmX2, X, (uint64_t)L.B * L.n, L.C, TILE, SLAB_CH, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&tmM, ws + L.w_M, (uint64_t)L.B * KP, L.C, KP, SLAB_CH, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  Params P;
  P.Rt = ws + L.w_Rt2; P.Ct = ws + L.w_Ct2; P.part = ws + L.w_PART;
  P.n = L.n; P.H = L.H; P.W = L.W; P.k = L.k; P.nsplit = L.nsplit_cen; P.tiles_per_image = L.n / TILE;
  P.nstages = nst;
  const int smem_bytes = CF::FIXED_BYTES + nst * SLAB_BYTES + 1024;
  if (const char* dbg = getenv("GF_DEBUG_PTR")) tc::set_debug_buffer(reinterpret_cast<unsigned int*>(strtoull(dbg, nullptr, 0)));
  auto kern = centroid_tc_kernel<KP, NS, NS2>;
  GF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  kern<<<dim3(L.nsplit_cen, L.B, NS / NS2), NUM_THREADS, smem_bytes, st>>>(tmX, tmX2, tmM, P);
  GF_LAUNCH_OK();
  return GF_OK;
}

}  // namespace tcc

bool tc_centroid_supported(const Layout& L, const gf_attn_desc* d) {
  static const bool disabled = getenv("GF_DISABLE_TC") != nullptr || getenv("GF_DISABLE_TC_CENTROID") != nullptr;
  if (disabled || (d->flags & GF_FLAG_FP32_EXACT)) return false;
  if (L.C != 64 && L.C != 128 && L.C != 256 && L.C != 512) return false;   // C = 512: two CTAs share the channels
  if (L.n % tcc::TILE != 0 || L.B > 65535) return false;
  const int limit = tc::device_smem_optin();
  const int ns = L.C / 32;
  if (L.KP == 16) return (ns == 2 ? tcc::stages_for<16, 2>(limit) : ns == 4 ? tcc::stages_for<16, 4>(limit) : ns == 8 ? tcc::stages_for<16, 8>(limit) : tcc::stages_for<16, 16>(limit)) >= 4;
  return (ns == 2 ? tcc::stages_for<32, 2>(limit) : ns == 4 ? tcc::stages_for<32, 4>(limit) : ns == 8 ? tcc::stages_for<32, 8>(limit) : tcc::stages_for<32, 16>(limit)) >= 4;
}

int centroid_pass_tc(const Layout& L, const gf_attn_desc* d, const float* X, float* ws, cudaStream_t st) {
  (void)d;
  const int ns = L.C / 32;
  if (L.KP == 16) return ns == 2 ? tcc::launch<16, 2>(L, X, ws, st) : ns == 4 ? tcc::launch<16, 4>(L, X, ws, st) : ns == 8 ? tcc::launch<16, 8>(L, X, ws, st) : tcc::launch<16, 16>(L, X, ws, st);
  return ns == 2 ? tcc::launch<32, 2>(L, X, ws, st) : ns == 4 ? tcc::launch<32, 4>(L, X, ws, st) : ns == 8 ? tcc::launch<32, 8>(L, X, ws, st) : tcc::launch<32, 16>(L, X, ws, st);
}

}  // namespace gf
