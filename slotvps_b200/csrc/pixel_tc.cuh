// tcgen05 / TMEM / TMA implementation of the pixel-side retriever contraction (kernel_path = 0).
//
// Precision: the <=1e-3 per-stage tolerance rules out single-pass bf16/tf32 operands,
// so every fp32 operand is split into fp16 hi + lo planes and each product is evaluated as
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (error ~2^-16, measured 1.7e-5 per stage).
//
// Data flow per level:
//   split_planes_kernel : x fp32 [T][256][P] (+pos) -> fp16 planes  xh, xl, (x+pos)h, (x+pos)l  [T*P][256]
//   stats_tc_kernel     : per 128-pixel tile, D_k = (x+pos) . Wk_c^T and D_v = x . Wv_c^T on the tensor
//                         pipe (TMA -> smem ring -> tcgen05.mma -> TMEM), epilogue reduces each pixel's
//                         256 outputs to the LayerNorm scale rs = rsqrt(mean((D+b)^2)+eps); D never
//                         leaves the SM.
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace slotvps {

struct TcStageOperands {
  __half* wplanes = nullptr;     // [4][256][256]: Wk_c hi, Wk_c lo, Wv_c hi, Wv_c lo (row-major [out][in])
  // Triangular form of the same statistics: Wc = Q R (Householder, fp64) => ||Wc u + bc||^2 = ||R u + Q^T bc||^2 with R upper
  // triangular, so the channels of k-subtile ks only reach the first 64 (ks + 1) outputs: 62.5 % of the MMA work.
  __half* rplanes = nullptr;     // [4][256][256]: R_k hi, R_k lo, R_v hi, R_v lo
  float* rbias = nullptr;        // [2][256]: Q_k^T bk_c, Q_v^T bv_c
  double* qr_scratch = nullptr;  // [2][257][256] fp64 working copies (weights + bias row) of the factorisation
};
struct TcWorkspace {
  // operand planes x hi, x lo, (x+pos) hi, (x+pos) lo, each stored as four 64-channel sub-planes [4 planes][4 ks][rows][64]:
  // a [128 px][64 ch] TMA box is then ONE contiguous 16 KB block (whole DRAM pages) instead of 128 rows at a 512-byte pitch
  __half* planes = nullptr;
  __half* planes_alt = nullptr;  // second plane set: levels alternate so the next level's fusion can overlap this level's stages
  float *ytab = nullptr, *xtab = nullptr;  // separable sine tables [128][h], [128][w] of the CURRENT level
  float *ytab_l[SLOTVPS_MAX_LEVELS] = {nullptr}, *xtab_l[SLOTVPS_MAX_LEVELS] = {nullptr};   // one pair per level (the side stream runs ahead)
  float *ytabT_l[SLOTVPS_MAX_LEVELS] = {nullptr};   // row tables transposed [h][128]: a pixel's 16 channels are one 64-byte broadcast load (fuse_tc)
  __half* gplanes = nullptr;     // [groups][T][2][104][256] folded query operand G, hi/lo (one copy per slot group)
  float2* ml = nullptr;          // [groups + 1][T*Pmax] per-pixel softmax (max, sum) of each slot group + combined (N > 104 only)
  long plane_rows = 0;                  // rows allocated per plane (T*Pmax)
};

// Separable position embedding (sine, or none): the kernels read only the x planes and add the position terms
// from small tables in their epilogues instead of streaming separate (x+pos) planes:
//   Wk_c (x+pos)      = Wk_c x + tky[row] + tkx[col]        tky [h][256], tkx [w][256]      (per stage)
//   (x+pos) . G_n     = x . G_n + pgy[row][n] + pgx[col][n]  pgy [T][h][112], pgx [T][w][112] (per stage, frame)
struct PosSep {
  int enabled = 0, w = 1, h = 1;
  const float *tky = nullptr, *tkx = nullptr;     // null with enabled: no position embedding at all
  const float *pgy = nullptr, *pgx = nullptr;
};

inline void tc_stage_layout(Arena& a, TcStageOperands* o) {
  o->wplanes = a.take<__half>((size_t)4 * C * C);
  o->rplanes = a.take<__half>((size_t)4 * C * C);
  o->rbias = a.take<float>(2 * C);
  o->qr_scratch = a.take<double>((size_t)2 * (C + 1) * C);
}
inline void tc_workspace_layout(Arena& a, const slotvps_head_desc* d, TcWorkspace* w) {
  long Pmax = 0;
  int hmax = 0, wmax = 0;
  for (int l = 0; l < d->n_levels; ++l) {
    Pmax = max(Pmax, (long)d->h[l] * d->w[l]);
    hmax = max(hmax, d->h[l]); wmax = max(wmax, d->w[l]);
  }
  w->plane_rows = (long)d->n_frames * Pmax;
  if (d->kernel_path == 0) {
    w->planes = a.take<__half>((size_t)4 * w->plane_rows * C);
    w->planes_alt = a.take<__half>((size_t)4 * w->plane_rows * C);
    for (int l = 0; l < SLOTVPS_MAX_LEVELS; ++l) { w->ytab_l[l] = a.take<float>((size_t)128 * hmax); w->xtab_l[l] = a.take<float>((size_t)128 * wmax); w->ytabT_l[l] = a.take<float>((size_t)128 * hmax); }
    w->ytab = w->ytab_l[0]; w->xtab = w->xtab_l[0];
    const int groups = (d->n_slots + 103) / 104;
    w->gplanes = a.take<__half>((size_t)groups * d->n_frames * 2 * 112 * C);
    if (groups > 1) w->ml = a.take<float2>((size_t)(groups + 1) * w->plane_rows);
  }
}
// shapes the tensor-core kernels serve; everything else runs on the fp32 path
inline bool tc_supported(const slotvps_head_desc* d, int level) { return d->h[level] * d->w[level] >= 128; }

// ---- operand preparation -----------------------------------------------------------------------------
// fp32 -> fp16 hi + lo (22 significant bits).  tcgen05 kind::f16 requires A and B to share one 16-bit
// format (mixing bf16 x fp16 raises an illegal-instruction fault on B200), and the softmax weights P
// need fp16's 11-bit mantissa, so every operand plane is fp16; values are clamped to the fp16 range.
__device__ __forceinline__ void split_bf16(float v, __half& hi, __half& lo) {
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
// two fp32 values -> packed fp16 hi pair + lo pair (hi + lo carries 22 bits; clamped to the fp16 range)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fminf(fmaxf(a, -65504.f), 65504.f); b = fminf(fmaxf(b, -65504.f), 65504.f);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}
__global__ void __launch_bounds__(256) weight_planes_kernel(const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                            __half* __restrict__ out) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= C * C) return;
  __half h, l;
  split_bf16(Wk[i], h, l); out[i] = h; out[C * C + i] = l;
  split_bf16(Wv[i], h, l); out[2 * C * C + i] = h; out[3 * C * C + i] = l;
}
// separable sine tables (position_encoding.py:236-256): channels [0,128) depend on the row, [128,256) on the column
__global__ void __launch_bounds__(256) pos_tab_kernel(float* __restrict__ ytab, float* __restrict__ xtab, int h, int w,
                                                      float* __restrict__ ytabT = nullptr /* optional [h][128] copy */) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 128 * (h + w)) return;
  const bool isy = i < 128 * h;
  int j = isy ? i : i - 128 * h;
  int n = isy ? h : w;
  int ci = j / n, r = j % n;
  float e = (float)(r + 1) / ((float)n + 1e-6f) * 6.283185307179586f;
  float a = e / powf(10000.f, (float)(2 * (ci / 2)) / 128.f);
  const float val = (ci & 1) ? cosf(a) : sinf(a);
  (isy ? ytab : xtab)[j] = val;
  if (isy && ytabT) ytabT[r * 128 + ci] = val;
}
// grid (ceil(P/32), T).  pos: tensor [256][P] per frame (pos_bs stride) | tables | none
constexpr int SPLIT_SMEM = 2 * 256 * 33 * (int)sizeof(float);
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, long x_bs, const float* __restrict__ pos, long pos_bs,
                                                           const float* __restrict__ ytab, const float* __restrict__ xtab,
                                                           __half* __restrict__ planes, long plane_rows, int P, int h, int w) {
  extern __shared__ float sp_smem[];
  float* xs = sp_smem;             // [256][33]
  float* ps = sp_smem + 256 * 33;  // [256][33]  x + pos
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = blockIdx.x * 32, t = blockIdx.y;
  const int p = p0 + lane;
  const bool pv = p < P;
  const float* __restrict__ X = x + (long)t * x_bs;
  const float* __restrict__ PZ = pos ? pos + (long)t * pos_bs : nullptr;
  const int row = pv ? p / w : 0, col = pv ? p % w : 0;
  for (int c = warp; c < C; c += 8) {
    float v = pv ? __ldg(X + (long)c * P + p) : 0.f;
    float q = 0.f;
    if (pv) {
      if (PZ) q = __ldg(PZ + (long)c * P + p);
      else if (ytab) q = c < 128 ? __ldg(ytab + c * h + row) : __ldg(xtab + (c - 128) * w + col);
    }
    xs[c * 33 + lane] = v;
    ps[c * 33 + lane] = v + q;
  }
  __syncthreads();
  __half2* out = reinterpret_cast<__half2*>(planes);
  const long plane_stride2 = 4 * plane_rows * 32;                    // one plane = 4 sub-planes of [rows][64]
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    int idx = tid + i * 256, pp = idx >> 7, cp = idx & 127;
    if (p0 + pp >= P) continue;
    const long o = ((long)(cp >> 5) * plane_rows + (long)t * P + p0 + pp) * 32 + (cp & 31);      // sub-plane ks = cp / 32, 32 half2 per row
    __half h0, l0, h1, l1;
    split_bf16(xs[(2 * cp) * 33 + pp], h0, l0); split_bf16(xs[(2 * cp + 1) * 33 + pp], h1, l1);
    out[o] = __halves2half2(h0, h1);
    out[plane_stride2 + o] = __halves2half2(l0, l1);
    split_bf16(ps[(2 * cp) * 33 + pp], h0, l0); split_bf16(ps[(2 * cp + 1) * 33 + pp], h1, l1);
    out[2 * plane_stride2 + o] = __halves2half2(h0, h1);
    out[3 * plane_stride2 + o] = __halves2half2(l0, l1);
  }
}

// ---- LayerNorm statistics of the key / value projections on the tensor pipe ---------------------------------
namespace stats {
constexpr int TILE_M = 128;                    // pixels per tile (TMEM lanes)
constexpr int KSUB = 64;                       // channels per k-subtile = one 128-byte swizzle row
constexpr int A_BYTES = TILE_M * 128;          // 16 KB  [128 px][64 ch] fp16
constexpr int B_BYTES = C * 128;               // 32 KB  [256 out][64 ch] fp16
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // A hi, A lo, B hi, B lo = 96 KB
constexpr int NSTAGE = 2;
constexpr int AUX_BYTES = 4096;                // barriers, tmem pointer, biases
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + AUX_BYTES + 1024;   // + alignment slack
constexpr int THREADS = 192;                   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr uint32_t IDESC = tc::make_idesc_f16(128, 256, 0, 0);
}  // namespace stats

// planes: tensor map over [4*plane_rows][256] fp16; wplanes: tensor map over [4*256][256] fp16
__global__ void __launch_bounds__(stats::THREADS, 1)
stats_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const float* __restrict__ bk_c, const float* __restrict__ bv_c, float* __restrict__ rs_k,
                float* __restrict__ rs_v, int P, int T, int plane_rows, int tiles_per_frame, const PosSep ps) {
  using namespace stats;
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);          // 1024-byte aligned ring base
  uint8_t* aux = smem + NSTAGE * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);                  // [NSTAGE]
  uint64_t* empty = full + NSTAGE;                                    // [NSTAGE]
  uint64_t* tfull = empty + NSTAGE;                                   // [2] accumulator ready
  uint64_t* tempty = tfull + 2;                                       // [2] accumulator drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias = reinterpret_cast<float*>(aux + 256);                  // [2][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = T * tiles_per_frame;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 128); }
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * C; i += THREADS) bias[i] = i < C ? bk_c[i] : bv_c[i - C];
  if (warp == 1) { tc::tmem_alloc(tmem_ptr, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int t = tile / tiles_per_frame, p0 = (tile % tiles_per_frame) * TILE_M;
        const int row = t * P + p0;
        for (int g = 0; g < 2; ++g) {                       // g = 0: keys (x+pos, Wk), g = 1: values (x, Wv)
          const int aq = (g == 0 && !ps.enabled) ? 2 : 0, bq = g == 0 ? 0 : 2;     // separable pos: keys read the x planes too
          for (int ks = 0; ks < C / KSUB; ++ks, ++it) {
            const int s = it % NSTAGE;
            tc::mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            tc::mbar_expect_tx(&full[s], STAGE_BYTES);
            tc::tma_load_2d(st, &tmap_x, 0, (aq * 4 + ks) * plane_rows + row, &full[s]);
            tc::tma_load_2d(st + A_BYTES, &tmap_x, 0, ((aq + 1) * 4 + ks) * plane_rows + row, &full[s]);
            tc::tma_load_2d(st + 2 * A_BYTES, &tmap_w, ks * KSUB, bq * C, &full[s]);
            tc::tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmap_w, ks * KSUB, (bq + 1) * C, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        for (int g = 0; g < 2; ++g) {
          tc::mbar_wait(&tempty[g], (ti & 1) ^ 1);          // epilogue has drained accumulator g of the previous tile
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + g * 256;
          for (int ks = 0; ks < C / KSUB; ++ks, ++it) {
            const int s = it % NSTAGE;
            tc::mbar_wait(&full[s], (it / NSTAGE) & 1);
            tc::tc_fence_after();
            const uint32_t a_hi = tc::smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
            const uint64_t dah = tc::make_smem_desc_sw128(a_hi, 16, 1024), dal = tc::make_smem_desc_sw128(a_lo, 16, 1024);
            const uint64_t dbh = tc::make_smem_desc_sw128(b_hi, 16, 1024), dbl = tc::make_smem_desc_sw128(b_lo, 16, 1024);
#pragma unroll
            for (int k = 0; k < KSUB / 16; ++k) {           // 16 channels = 32 bytes = +2 in the address field
              tc::umma_bf16(d_tmem, dah + 2 * k, dbh + 2 * k, IDESC, (ks | k) != 0);
              tc::umma_bf16(d_tmem, dal + 2 * k, dbh + 2 * k, IDESC, 1);
              tc::umma_bf16(d_tmem, dah + 2 * k, dbl + 2 * k, IDESC, 1);
            }
            tc::umma_commit(&empty[s]);                     // smem slot reusable once these MMAs retire
          }
          tc::umma_commit(&tfull[g]);                       // accumulator g complete
        }
      }
    }
  } else {
    // ===================== epilogue: per-pixel sum of squares over the 256 outputs =====================
    const int q = warp & 3;                                 // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;                            // pixel row inside the tile
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int t = tile / tiles_per_frame, p = (tile % tiles_per_frame) * TILE_M + r;
      for (int g = 0; g < 2; ++g) {
        tc::mbar_wait(&tfull[g], ti & 1);
        tc::tc_fence_after();
        const float* bg = bias + g * C;
        const bool add_pos = g == 0 && ps.tky != nullptr && p < P;
        const float4* ty4 = add_pos ? reinterpret_cast<const float4*>(ps.tky + (long)(p / ps.w) * C) : nullptr;
        const float4* tx4 = add_pos ? reinterpret_cast<const float4*>(ps.tkx + (long)(p % ps.w) * C) : nullptr;
        float ss = 0.f;
#pragma unroll 1
        for (int j = 0; j < C / 32; ++j) {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + g * 256 + j * 32, v);
          tc::tmem_ld_wait();
          if (add_pos) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 a = __ldg(ty4 + j * 8 + c), b = __ldg(tx4 + j * 8 + c);
              v[4 * c] += a.x + b.x; v[4 * c + 1] += a.y + b.y; v[4 * c + 2] += a.z + b.z; v[4 * c + 3] += a.w + b.w;
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) { float d = v[c] + bg[j * 32 + c]; ss = fmaf(d, d, ss); }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&tempty[g]);
        if (p < P) (g == 0 ? rs_k : rs_v)[(long)t * P + p] = rsqrtf(ss * (1.f / C) + LN_EPS);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// ---- the same statistics on CTA pairs (cta_group::2) ----------------------------------------------------------
// stats_tc_kernel ingests 96 KB per pipeline stage (32 KB of planes + 64 KB of weight planes) for ~1536 MMA cycles,
// i.e. it sits at the ~64 B/clk per-SM L2->shared-memory limit.  Here a cluster of two CTAs runs one M = 256 MMA
// over two pixel tiles and each CTA stages only HALF of the weight tile (128 of the 256 output rows): 64 KB per
// stage and SM, which also makes room for a third pipeline stage.
namespace stats2 {
constexpr int TILE_M = 128, KSUB = 64;
constexpr int A_BYTES = TILE_M * 128;          // 16 KB  [128 px][64 ch] fp16
constexpr int B_BYTES = 128 * 128;             // 16 KB  [128 out][64 ch] fp16 : this CTA's half of the weight tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;    // 64 KB
constexpr int NSTAGE = 3;
constexpr int AUX_BYTES = 4096;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + AUX_BYTES + 1024;
constexpr int THREADS = 192;
constexpr uint32_t IDESC = tc::make_idesc_f16(256, 256, 0, 0);
}  // namespace stats2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(stats2::THREADS, 1)
stats_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const float* __restrict__ bk_c, const float* __restrict__ bv_c, float* __restrict__ rs_k,
                 float* __restrict__ rs_v, int P, int T, int plane_rows, int tiles_per_frame, const PosSep ps) {
  using namespace stats2;
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  uint8_t* aux = smem + NSTAGE * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);                  // [NSTAGE] used on the leader
  uint64_t* empty = full + NSTAGE;                                    // [NSTAGE] in both CTAs (multicast commit)
  uint64_t* tfull = empty + NSTAGE;                                   // [2] in both CTAs (multicast commit)
  uint64_t* tempty = tfull + 2;                                       // [2] on the leader: 256 arrivals (both CTAs' epilogues)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias = reinterpret_cast<float*>(aux + 256);                  // [2][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles = T * tiles_per_frame;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_iter = ((n_tiles + 1) / 2 - pair + n_pairs - 1) / n_pairs;      // identical for both CTAs of the pair

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 256); }
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * C; i += THREADS) bias[i] = i < C ? bk_c[i] : bv_c[i - C];
  if (warp == 1) { tc::tmem_alloc2(tmem_ptr, 512); tc::tmem_relinquish2(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();                                                 // the peer's barriers exist before anything targets them
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own pixel tile + own half of the weight tile =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int i = 0; i < n_iter; ++i) {
        const int tile = (pair + i * n_pairs) * 2 + (int)rank;        // may be n_tiles (odd tail): rows past the data, results dropped
        const int t = tile / tiles_per_frame, p0 = (tile % tiles_per_frame) * TILE_M;
        const int row = t * P + p0;
        for (int g = 0; g < 2; ++g) {
          const int aq = (g == 0 && !ps.enabled) ? 2 : 0, bq = g == 0 ? 0 : 2;
          for (int ks = 0; ks < C / KSUB; ++ks, ++it) {
            const int s = it % NSTAGE;
            tc::mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            if (leader) tc::mbar_expect_tx(&full[s], 2 * STAGE_BYTES);           // bytes of both CTAs land on this barrier
            tc::tma_load_2d_pair(st, &tmap_x, 0, (aq * 4 + ks) * plane_rows + row, &full[s]);
            tc::tma_load_2d_pair(st + A_BYTES, &tmap_x, 0, ((aq + 1) * 4 + ks) * plane_rows + row, &full[s]);
            tc::tma_load_2d_pair(st + 2 * A_BYTES, &tmap_w, ks * KSUB, bq * C + (int)rank * 128, &full[s]);
            tc::tma_load_2d_pair(st + 2 * A_BYTES + B_BYTES, &tmap_w, ks * KSUB, (bq + 1) * C + (int)rank * 128, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only): M = 256 over both CTAs =====================
    if (lane == 0 && leader) {
      uint32_t it = 0;
      for (int i = 0; i < n_iter; ++i) {
        for (int g = 0; g < 2; ++g) {
          tc::mbar_wait(&tempty[g], (i & 1) ^ 1);
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + g * 256;
          for (int ks = 0; ks < C / KSUB; ++ks, ++it) {
            const int s = it % NSTAGE;
            tc::mbar_wait(&full[s], (it / NSTAGE) & 1);
            tc::tc_fence_after();
            const uint32_t a_hi = tc::smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
            const uint64_t dah = tc::make_smem_desc_sw128(a_hi, 16, 1024), dal = tc::make_smem_desc_sw128(a_lo, 16, 1024);
            const uint64_t dbh = tc::make_smem_desc_sw128(b_hi, 16, 1024), dbl = tc::make_smem_desc_sw128(b_lo, 16, 1024);
#pragma unroll
            for (int k = 0; k < KSUB / 16; ++k) {
              tc::umma2_f16(d_tmem, dah + 2 * k, dbh + 2 * k, IDESC, (ks | k) != 0);
              tc::umma2_f16(d_tmem, dal + 2 * k, dbh + 2 * k, IDESC, 1);
              tc::umma2_f16(d_tmem, dah + 2 * k, dbl + 2 * k, IDESC, 1);
            }
            tc::umma2_commit_multicast(&empty[s], 3);       // slot s is free in both CTAs
          }
          tc::umma2_commit_multicast(&tfull[g], 3);         // accumulator g complete in both CTAs' TMEM
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): each drains its own 128 accumulator lanes =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    for (int i = 0; i < n_iter; ++i) {
      const int tile = (pair + i * n_pairs) * 2 + (int)rank;
      const int t = tile / tiles_per_frame, p = (tile % tiles_per_frame) * TILE_M + r;
      const bool valid = tile < n_tiles && p < P;
      for (int g = 0; g < 2; ++g) {
        tc::mbar_wait(&tfull[g], i & 1);
        tc::tc_fence_after();
        const float* bg = bias + g * C;
        const bool add_pos = g == 0 && ps.tky != nullptr && valid;
        const float4* ty4 = add_pos ? reinterpret_cast<const float4*>(ps.tky + (long)(p / ps.w) * C) : nullptr;
        const float4* tx4 = add_pos ? reinterpret_cast<const float4*>(ps.tkx + (long)(p % ps.w) * C) : nullptr;
        float ss = 0.f;
#pragma unroll 1
        for (int j = 0; j < C / 32; ++j) {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + g * 256 + j * 32, v);
          tc::tmem_ld_wait();
          if (add_pos) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 a = __ldg(ty4 + j * 8 + c), b = __ldg(tx4 + j * 8 + c);
              v[4 * c] += a.x + b.x; v[4 * c + 1] += a.y + b.y; v[4 * c + 2] += a.z + b.z; v[4 * c + 3] += a.w + b.w;
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) { float d = v[c] + bg[j * 32 + c]; ss = fmaf(d, d, ss); }
        }
        tc::tc_fence_before();
        tc::mbar_arrive_remote(&tempty[g], 0);              // the leader's MMA waits for the epilogues of both CTAs
        if (valid) (g == 0 ? rs_k : rs_v)[(long)t * P + p] = rsqrtf(ss * (1.f / C) + LN_EPS);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();                                                 // neither CTA leaves while the other may still signal it
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc2(tmem_base, 512); }
}

// ---- triangular statistics -------------------------------------------------------------------------------------------
// Householder QR of the centred projection (rows = outputs, columns = input channels) in fp64, one CTA per matrix:
// A <- H_255 ... H_0 A = R (upper triangular), b <- Q^T b.  Thread c owns column c (and thread 0 the bias, stored as
// column-vector row 256 of the scratch: scratch[i][c] for i < 256 is A, scratch[256][i] is b_i).
// blockIdx.x = 0: keys (Wk_c, bk_c), 1: values (Wv_c, bv_c).
__global__ void __launch_bounds__(256) qr_planes_kernel(const float* __restrict__ Wk, const float* __restrict__ bk, const float* __restrict__ Wv,
                                                        const float* __restrict__ bv, double* __restrict__ scratch, __half* __restrict__ planes,
                                                        float* __restrict__ rbias) {
  const int which = blockIdx.x, c = threadIdx.x;
  const float* W = which ? Wv : Wk;
  const float* b = which ? bv : bk;
  double* A = scratch + (size_t)which * (C + 1) * C;
  double* bb = A + (size_t)C * C;
  __shared__ double sv[C];
  __shared__ double red[8];
  __shared__ double s_beta;
  for (int i = 0; i < C; ++i) A[i * C + c] = (double)W[i * C + c];
  bb[c] = (double)b[c];
  __syncthreads();
  for (int j = 0; j < C - 1; ++j) {
    // v = x + sign(x_0) ||x|| e_0 for x = A[j:, j]
    const double xj = c >= j ? A[c * C + j] : 0.0;          // thread c holds row c of column j
    double part = xj * xj;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((c & 31) == 0) red[c >> 5] = part;
    __syncthreads();
    double nrm2 = 0.0;
    for (int i = 0; i < 8; ++i) nrm2 += red[i];
    const double nrm = sqrt(nrm2);
    const double x0 = A[j * C + j];
    const double alpha = x0 >= 0.0 ? -nrm : nrm;
    sv[c] = c < j ? 0.0 : (c == j ? xj - alpha : xj);
    if (c == 0) { const double v0 = x0 - alpha; s_beta = (nrm2 - x0 * x0 + v0 * v0); }     // ||v||^2
    __syncthreads();
    const double vv = s_beta;
    if (vv > 0.0) {
      // column c (>= j) of A, and the bias through thread j's spare time: thread 255 handles b when c == 255
      if (c >= j) {
        double dot = 0.0;
        for (int i = j; i < C; ++i) dot += sv[i] * A[i * C + c];
        const double f = 2.0 * dot / vv;
        for (int i = j; i < C; ++i) A[i * C + c] -= f * sv[i];
      }
      if (c == 0) {
        double dot = 0.0;
        for (int i = j; i < C; ++i) dot += sv[i] * bb[i];
        const double f = 2.0 * dot / vv;
        for (int i = j; i < C; ++i) bb[i] -= f * sv[i];
      }
    }
    __syncthreads();
  }
  // R (entries below the diagonal are exactly zero by construction) -> fp16 hi / lo planes, Q^T b -> fp32.  Plane row =
  // accumulator column of the CTA-pair kernel: output o of "ring" o / 64 sits 32 (ring + 1) columns left of the centre
  // (first half of the ring, rows of CTA 0) or 32 ring columns right of it (second half, CTA 1), so the outputs that are
  // active for k-subtile ks -- o < 64 (ks + 1) -- occupy the contiguous, centred column range [32 (3 - ks), 256 - 32 (3 - ks)).
  __half* hi = planes + (size_t)(2 * which) * C * C;
  __half* lo = hi + (size_t)C * C;
  for (int o = 0; o < C; ++o) {
    const int ring = o >> 6, off = o & 63;
    const int col = off < 32 ? 128 - 32 * (ring + 1) + off : 128 + 32 * ring + (off - 32);
    double v = c >= o ? A[o * C + c] : 0.0;
    v = fmin(fmax(v, -65504.0), 65504.0);
    const __half h = __double2half(v);
    hi[col * C + c] = h;
    lo[col * C + c] = __double2half(v - (double)__half2float(h));
  }
  {
    const int ring = c >> 6, off = c & 63;
    const int col = off < 32 ? 128 - 32 * (ring + 1) + off : 128 + 32 * ring + (off - 32);
    rbias[which * C + col] = (float)bb[c];
  }
}

namespace stats3 {
constexpr int TILE_M = 128, KSUB = 64;
constexpr int A_BYTES = TILE_M * 128;          // 16 KB
constexpr int B_ROWS = 32;                     // weight rows per TMA box
constexpr int B_BOX = B_ROWS * 128;            // 4 KB
constexpr int B_BYTES = 128 * 128;             // up to 128 rows of one plane per CTA (half of the active outputs)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;    // 64 KB (filled 40 / 48 / 56 / 64 KB)
constexpr int NSTAGE = 3;
constexpr int AUX_BYTES = 4096;
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + AUX_BYTES + 1024;
constexpr int THREADS = 192;
}  // namespace stats3

// The statistics from the triangular factors on CTA pairs (cta_group::2, M = 256 = two pixel tiles).  k-subtiles are visited
// from the LAST to the first: the first MMA of a tile (N = 256) initialises every accumulator column, the later, narrower
// ones (N = 192, 128, 64 at column offset 32, 64, 96) only add to theirs.  Each CTA stages its half of the active weight rows.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(stats3::THREADS, 1)
stats_tri_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_r0, const __grid_constant__ CUtensorMap tmap_r1,
                 const __grid_constant__ CUtensorMap tmap_r2, const __grid_constant__ CUtensorMap tmap_r3,      // boxes of 32, 64, 96, 128 weight rows
                 const float* __restrict__ rbias, float* __restrict__ rs_k, float* __restrict__ rs_v, int P, int T, int plane_rows,
                 int tiles_per_frame, int x_planes_only, int products) {
  using namespace stats3;
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  uint8_t* aux = smem + NSTAGE * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);                  // [NSTAGE] used on the leader
  uint64_t* empty = full + NSTAGE;                                    // [NSTAGE] in both CTAs (multicast commit)
  uint64_t* tfull = empty + NSTAGE;                                   // [2] in both CTAs (multicast commit)
  uint64_t* tempty = tfull + 2;                                       // [2] on the leader: 256 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* bias = reinterpret_cast<float*>(aux + 256);                  // [2][256] in accumulator-column order
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
  const uint32_t rank = tc::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles = T * tiles_per_frame;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_iter = ((n_tiles + 1) / 2 - pair + n_pairs - 1) / n_pairs;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_r0); tc::tma_prefetch_desc(&tmap_r1); tc::tma_prefetch_desc(&tmap_r2); tc::tma_prefetch_desc(&tmap_r3);
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 256); }
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * C; i += THREADS) bias[i] = rbias[i];
  if (warp == 1) { tc::tmem_alloc2(tmem_ptr, 512); tc::tmem_relinquish2(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int i = 0; i < n_iter; ++i) {
        const int tile = (pair + i * n_pairs) * 2 + (int)rank;
        const int t = tile / tiles_per_frame, p0 = (tile % tiles_per_frame) * TILE_M;
        const int row = t * P + p0;
        for (int g = 0; g < 2; ++g) {
          const int aq = (g == 0 && !x_planes_only) ? 2 : 0, bq = g == 0 ? 0 : 2;
          for (int ks = C / KSUB - 1; ks >= 0; --ks, ++it) {
            const int s = it % NSTAGE, nb = ks + 1;                  // nb boxes of 32 weight rows per plane and CTA
            const int r0 = rank == 0 ? 32 * (3 - ks) : 128;          // first plane row (= accumulator column) of this CTA's active half
            tc::mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            if (leader) tc::mbar_expect_tx(&full[s], 2 * (2 * A_BYTES + 2 * nb * B_BOX));
            tc::tma_load_2d_pair(st, &tmap_x, 0, (aq * 4 + ks) * plane_rows + row, &full[s]);
            tc::tma_load_2d_pair(st + A_BYTES, &tmap_x, 0, ((aq + 1) * 4 + ks) * plane_rows + row, &full[s]);
            const CUtensorMap* mr = ks == 0 ? &tmap_r0 : ks == 1 ? &tmap_r1 : ks == 2 ? &tmap_r2 : &tmap_r3;     // one box of 32 nb rows per plane
            tc::tma_load_2d_pair(st + 2 * A_BYTES, mr, ks * KSUB, bq * C + r0, &full[s]);
            tc::tma_load_2d_pair(st + 2 * A_BYTES + B_BYTES, mr, ks * KSUB, (bq + 1) * C + r0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {                                             // all lanes of the leader CTA's warp: warp-uniform issue loop (tc::elect_one)
      const bool el = tc::elect_one();
      uint32_t it = 0;
      for (int i = 0; i < n_iter; ++i) {
        for (int g = 0; g < 2; ++g) {
          tc::mbar_wait(&tempty[g], (i & 1) ^ 1);
          tc::tc_fence_after();
          for (int ks = C / KSUB - 1; ks >= 0; --ks, ++it) {
            const int s = it % NSTAGE;
            const uint32_t idesc = tc::make_idesc_f16(256, 64 * (ks + 1), 0, 0);
            const uint32_t d_tmem = tmem_base + g * 256 + 32 * (3 - ks);
            tc::mbar_wait(&full[s], (it / NSTAGE) & 1);
            tc::tc_fence_after();
            const uint32_t a_hi = tc::smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
            const uint64_t dah = tc::make_smem_desc_sw128(a_hi, 16, 1024), dal = tc::make_smem_desc_sw128(a_lo, 16, 1024);
            const uint64_t dbh = tc::make_smem_desc_sw128(b_hi, 16, 1024), dbl = tc::make_smem_desc_sw128(b_lo, 16, 1024);
            if (el) {
#pragma unroll
              for (int k = 0; k < KSUB / 16; ++k) {
                tc::umma2_f16(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, (ks != C / KSUB - 1 || k != 0) ? 1u : 0u);
                if (products >= 2) tc::umma2_f16(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, 1);      // experiment switch; 3 = shipped
                if (products >= 3) tc::umma2_f16(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1);
              }
              tc::umma2_commit_multicast(&empty[s], 3);
            }
          }
          if (el) tc::umma2_commit_multicast(&tfull[g], 3);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    for (int i = 0; i < n_iter; ++i) {
      const int tile = (pair + i * n_pairs) * 2 + (int)rank;
      const int t = tile / tiles_per_frame, p = (tile % tiles_per_frame) * TILE_M + r;
      const bool valid = tile < n_tiles && p < P;
      for (int g = 0; g < 2; ++g) {
        tc::mbar_wait(&tfull[g], i & 1);
        tc::tc_fence_after();
        const float* bg = bias + g * C;
        float ss = 0.f;
#pragma unroll 1
        for (int j = 0; j < C / 32; ++j) {
          float v[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + g * 256 + j * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) { float d = v[c] + bg[j * 32 + c]; ss = fmaf(d, d, ss); }
        }
        tc::tc_fence_before();
        tc::mbar_arrive_remote(&tempty[g], 0);
        if (valid) (g == 0 ? rs_k : rs_v)[(long)t * P + p] = rsqrtf(ss * (1.f / C) + LN_EPS);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc2(tmem_base, 512); }
}

// ---- host side ----------------------------------------------------------------------------------------------
inline int tc_prepare_stage(const slotvps_stage_params& sp, const float* Wk_c, const float* bk_c, const float* Wv_c,
                            const float* bv_c, TcStageOperands& o, cudaStream_t s) {
  (void)sp; (void)bk_c; (void)bv_c;
  weight_planes_kernel<<<ceil_div(C * C, 256), 256, 0, s>>>(Wk_c, Wv_c, o.wplanes);
  SV_CHECK_LAUNCH("weight_planes");
  qr_planes_kernel<<<2, 256, 0, s>>>(Wk_c, bk_c, Wv_c, bv_c, o.qr_scratch, o.rplanes, o.rbias);
  SV_CHECK_LAUNCH("qr_planes");
  return SLOTVPS_OK;
}

// x fp32 [T][256][P] (+ pos tensor | sine tables | nothing) -> bf16 operand planes of this level
inline int tc_split_level(const float* x, long x_bs, const float* pos, long pos_bs, bool sine, const TcWorkspace& ws, int T, int h, int w,
                          cudaStream_t s) {
  const int P = h * w;
  if ((long)T * P > ws.plane_rows || !ws.planes) return fail(SLOTVPS_EWORKSPACE, "tensor-core plane workspace too small%s%s");
  // planes of a level are packed with stride T*P rows, so a tile's rows past the end of a frame/plane are
  // rows of the next frame/plane (finite) or beyond the tensor (TMA zero fill) -- never stale memory
  const long rows = (long)T * P;
  SV_TRY(ensure_dyn_smem((const void*)split_planes_kernel, SPLIT_SMEM));
  if (sine && !pos) {
    pos_tab_kernel<<<ceil_div(128 * (h + w), 256), 256, 0, s>>>(ws.ytab, ws.xtab, h, w);
    SV_CHECK_LAUNCH("pos_tab");
  }
  split_planes_kernel<<<dim3(ceil_div(P, 32), T), 256, SPLIT_SMEM, s>>>(x, x_bs, pos, pos_bs, (sine && !pos) ? ws.ytab : nullptr,
                                                                         (sine && !pos) ? ws.xtab : nullptr, ws.planes, rows, P, h, w);
  SV_CHECK_LAUNCH("split_planes");
  return SLOTVPS_OK;
}

// rs_k, rs_v [T][P] from the planes of this level and the stage's weight planes
inline int tc_stats(const TcStageOperands& ops, const TcWorkspace& ws, const float* bk_c, const float* bv_c, float* rs_k, float* rs_v,
                    int T, int P, cudaStream_t s, int max_ctas = 148, const PosSep& ps = PosSep()) {
  CUtensorMap mx, mw;
  const long rows = (long)T * P;
  // only the planes that were written are inside the tensor map: a tail tile's rows past the last plane are then
  // zero-filled by TMA instead of reading stale memory (NaN bit patterns there corrupt the MMA even against zero weights)
  SV_TRY(tc::make_tmap_h16_sw128(&mx, ws.planes, (uint64_t)(ps.enabled ? 2 : 4) * 4 * rows, 64, stats::TILE_M));
  SV_TRY(tc::make_tmap_h16_sw128(&mw, ops.wplanes, (uint64_t)4 * C, C, C));
  SV_TRY(ensure_dyn_smem((const void*)stats_tc_kernel, stats::SMEM_BYTES));
  const int tiles_per_frame = ceil_div(P, stats::TILE_M);
  const int n_tiles = T * tiles_per_frame;
  static const int use_tri = getenv("SLOTVPS_STATS_TRI") ? atoi(getenv("SLOTVPS_STATS_TRI")) : 1;
  if (use_tri && ps.tky == nullptr && ops.rplanes != nullptr && n_tiles >= 2 && max_ctas >= 2) {      // triangular factors on CTA pairs
    CUtensorMap mr[4];                                                                                // (no separable position tables in this form)
    for (int i = 0; i < 4; ++i) SV_TRY(tc::make_tmap_h16_sw128(&mr[i], ops.rplanes, (uint64_t)4 * C, C, stats3::B_ROWS * (i + 1)));
    SV_TRY(ensure_dyn_smem((const void*)stats_tri_kernel, stats3::SMEM_BYTES));
    int grid3 = 2 * ((n_tiles + 1) / 2);
    if (grid3 > (max_ctas & ~1)) grid3 = max_ctas & ~1;
    g_prof_grid = grid3;
    // measurement switch only: 1 = hi.hi, 2 = hi.hi + lo_x.hi_w, 3 (default, shipped) = + hi_x.lo_w; read per call so one process can compare
    const char* pe = getenv("SLOTVPS_STATS_PRODUCTS");
    const int products = pe && atoi(pe) >= 1 && atoi(pe) <= 3 ? atoi(pe) : 3;
    stats_tri_kernel<<<grid3, stats3::THREADS, stats3::SMEM_BYTES, s>>>(mx, mr[0], mr[1], mr[2], mr[3], ops.rbias, rs_k, rs_v, P, T, (int)rows, tiles_per_frame,
                                                                        ps.enabled, products);
    SV_CHECK_LAUNCH("stats_tc");
    return SLOTVPS_OK;
  }
  static const int use_pairs = getenv("SLOTVPS_STATS_PAIRS") ? atoi(getenv("SLOTVPS_STATS_PAIRS")) : 1;   // CTA pairs by default
  if (use_pairs && n_tiles >= 2 && max_ctas >= 2) {
    CUtensorMap mw2;
    SV_TRY(tc::make_tmap_h16_sw128(&mw2, ops.wplanes, (uint64_t)4 * C, C, 128));
    SV_TRY(ensure_dyn_smem((const void*)stats_tc2_kernel, stats2::SMEM_BYTES));
    int grid2 = 2 * ((n_tiles + 1) / 2);
    if (grid2 > (max_ctas & ~1)) grid2 = max_ctas & ~1;
    g_prof_grid = grid2;
    stats_tc2_kernel<<<grid2, stats2::THREADS, stats2::SMEM_BYTES, s>>>(mx, mw2, bk_c, bv_c, rs_k, rs_v, P, T, (int)rows, tiles_per_frame, ps);
    SV_CHECK_LAUNCH("stats_tc");
    return SLOTVPS_OK;
  }
  const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
  g_prof_grid = grid;
  stats_tc_kernel<<<grid, stats::THREADS, stats::SMEM_BYTES, s>>>(mx, mw, bk_c, bv_c, rs_k, rs_v, P, T, (int)rows, tiles_per_frame, ps);
  SV_CHECK_LAUNCH("stats_tc");
  return SLOTVPS_OK;
}


// =====================================================================================================
// Fused slot-axis-softmax attention on the tensor pipe (replaces slot_attn_fp32_kernel):
//   S[p,n]  = (x+pos)_p . G_n                      tcgen05, M = 128 pixels (TMEM lanes), N = 112 slots
//   A[p,:]  = softmax_n(rs_k[p] (S + g0) + g1)     thread-local: one pixel per thread, slots in registers
//   P[n,p]  = A[p,n] * rs_v[p]            (fp16)   written to smem as the K-major B operand of the next MMA
//   Z^T[c,n] += sum_p x[p,c] P[n,p]                tcgen05, M = 128 channels x 2, N = 112, K = 128 pixels
//   aux[n,:] += sum_p P[n,p] (1, sigma_v[p])       tcgen05, N = 16  -> a1[n], a0[n]
// The N x P attention matrix never leaves the SM.  x / x+pos are the fp16 hi/lo planes of the level
// (3-product split for S; 2-product x_hi + x_lo against the fp16 P for Z, whose ~2^-12 rounding is averaged
// over the pixel sum); accumulation is fp32 in TMEM, Z^T stays resident in TMEM across the CTA's tiles.
namespace attn {
constexpr int TILE_M = 128;
constexpr int NROW = 104;                         // slot rows held in shared memory (N <= 104)
constexpr int NPAD = 112;                         // UMMA N (multiple of 16 for M = 128); rows 104..111 over-read finite data
constexpr int SLOT_BYTES = 16384;                 // ring slot: [128 px][64 ch] fp16, 128B swizzle
constexpr int NSLOT = 4;
constexpr int G_SUB = NROW * 128;                 // 13312: [104 slots][64 ch] fp16
constexpr int G_BYTES = 2 * 4 * G_SUB;            // hi, lo planes x 4 k-subtiles = 106496
constexpr int P_SUB = NROW * 128;                 // [104 slots][64 px] fp16
constexpr int P_PLANE = 2 * P_SUB;                // two 64-pixel halves
constexpr int P_BYTES = 2 * P_PLANE;              // hi, lo = 53248
constexpr int AUX_SUB = 16 * 128;                 // [16 pseudo-channels][64 px] fp16
constexpr int AUXT_BYTES = 2 * AUX_SUB;
constexpr int MISC_BYTES = 3072;                  // barriers [0,128), (g0, g1) [128,1024), softmax exchange [1024,3072)
constexpr int OFF_G = NSLOT * SLOT_BYTES;
constexpr int OFF_P = OFF_G + G_BYTES;
constexpr int OFF_AUX = OFF_P + P_BYTES;
constexpr int OFF_MISC = OFF_AUX + AUXT_BYTES;
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES;             // == 232448, the sm_100 per-block maximum: no alignment slack, the
                                                              // dynamic shared-memory array is declared __align__(1024) (checked at run time)
constexpr int THREADS = 320;                      // warp 0 TMA, warp 1 MMA, warps 2..9 softmax: TWO threads per pixel
constexpr int NSPLIT = 48;                        // slots [0,48) belong to the pixel's first thread, [48,112) to the second
constexpr uint32_t IDESC_S = tc::make_idesc_f16(128, NPAD, 0, 0);
// Z^T: A = x tile, MN-major fp16; B = P, K-major fp16
constexpr uint32_t IDESC_Z = tc::make_idesc_f16(128, NPAD, 1, 0);
// aux: A = P (K-major), B = aux tile (K-major), N = 16
constexpr uint32_t IDESC_AUX = tc::make_idesc_f16(128, 16, 0, 0);
constexpr int TM_S = 0;                            // two S buffers of 128 columns
constexpr int TM_Z = 256;                          // Z^T: 2 x 112 columns
constexpr int TM_AUX = 480;                        // 16 columns
// P = A * rs_v is scaled by 2^7 before the fp16 hi/lo split: fp16 has only 5 exponent bits, so the lo plane of a value
// below ~0.06 is subnormal (absolute step 6e-8) and softmax weights of ~1/N lose ~2 decimal digits (measured: 5.4e-6
// instead of 2.5e-6 per stage).  rs_v <= 1/sqrt(eps) = 316, so the scaled value stays below the fp16 maximum.
constexpr float PSCALE = 128.f, PSCALE_INV = 1.f / 128.f;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(OFF_G % 1024 == 0 && G_SUB % 1024 == 0 && OFF_P % 1024 == 0 && P_SUB % 1024 == 0 && OFF_AUX % 1024 == 0, "swizzle atoms");
}  // namespace attn

__global__ void __launch_bounds__(256) g_planes_kernel(const float* __restrict__ G, __half* __restrict__ out, int N, int T, int n0 = 0, int Ntot = 0) {
  // out [T][2][NROW][256] for slots n0 .. n0+N-1 of the Ntot rows of G per frame; rows >= N are zero
  if (Ntot == 0) Ntot = N;
  long i = (long)blockIdx.x * 256 + threadIdx.x;
  long total = (long)T * attn::NROW * C;
  if (i >= total) return;
  int c = (int)(i % C), n = (int)((i / C) % attn::NROW), t = (int)(i / ((long)C * attn::NROW));
  __half h = __float2half_rn(0.f), l = h;
  if (n < N) split_bf16(G[((long)t * Ntot + n0 + n) * C + c], h, l);
  long o = ((long)t * 2 * attn::NROW + n) * C + c;
  out[o] = h;
  out[o + (long)attn::NROW * C] = l;
}

// Slot groups (N > 104): the softmax runs over ALL slots of a pixel, so the kernel has three modes.
//   MODE 0  one group holds every slot: S, softmax, Z in one pass (the shipped N = 100 case).
//   MODE 1  denominator pass of one group: S only; writes the group's per-pixel (max, sum of exp) to grp.ml_out.
//   MODE 2  accumulation pass of one group: softmax weights exp(S - M) / L with the COMBINED (M, L) of grp.ml_in.
struct SlotGroup {
  int n0 = 0, n_total = 0;              // first slot of the group, slots per frame overall (N of the kernel = slots in the group)
  const float2* ml_in = nullptr;        // [T*P] combined (M, L)
  float2* ml_out = nullptr;             // [T*P] this group's (m, l)
};

template <int MODE>
__global__ void __launch_bounds__(attn::THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
               const float* __restrict__ g0, const float* __restrict__ g1, const float* __restrict__ rs_k,
               const float* __restrict__ rs_v, float* __restrict__ Zpart, float* __restrict__ a0part,
               float* __restrict__ a1part, const __half* __restrict__ planes, int N, int P, int T, int plane_rows, int tiles_per_frame, int dbg, const PosSep ps,
               const SlotGroup grp) {
  using namespace attn;
  const int Ntot = MODE == 0 ? N : grp.n_total, n0 = MODE == 0 ? 0 : grp.n0;
  extern __shared__ __align__(1024) uint8_t attn_smem[];
  uint8_t* smem = attn_smem;
  if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();          // the swizzled tiles need 1024-byte alignment and there is no spare byte to fix it up
  uint8_t* misc = smem + OFF_MISC;
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);      // [NSLOT]
  uint64_t* empty = full + NSLOT;                          // [NSLOT]
  uint64_t* sfull = empty + NSLOT;                         // [2]
  uint64_t* sempty = sfull + 2;                            // [2]
  uint64_t* pfull = sempty + 2;                            // [1]
  uint64_t* pempty = pfull + 1;                            // [1]
  uint64_t* gfull = pempty + 1;                            // [1] phase 0: G operand landed (TMA); phase 1: the Z / aux accumulators are complete
  uint64_t* zfull = gfull;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(gfull + 1);      // byte 120 of the 128-byte barrier block
  float2* gc = reinterpret_cast<float2*>(misc + 128);      // [NPAD] (g0, g1)
  float* xch = reinterpret_cast<float*>(misc + 1024);      // [2 kinds][2 halves][128] softmax max / sum exchange between a pixel's two threads

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
  const int chunk = blockIdx.x, chunks = gridDim.x, t = blockIdx.y;
  const int n_my = (tiles_per_frame - chunk + chunks - 1) / chunks;       // tiles chunk, chunk+chunks, ...

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_g);
    for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&sfull[i], 1); tc::mbar_init(&sempty[i], 256); }
    tc::mbar_init(pfull, 256); tc::mbar_init(pempty, 1); tc::mbar_init(gfull, 1);
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NPAD; i += THREADS)
    gc[i] = i < N ? make_float2(g0[(long)t * Ntot + n0 + i], g1[(long)t * Ntot + n0 + i]) : make_float2(0.f, 0.f);
  // aux tile: row 0 = 1.0 (-> a1), rows 1 / 2 = sigma_v per pixel as fp16 hi / lo (rewritten every tile), rows 3..15 = 0.
  // Row r lives at byte r*128 of each 64-pixel half; 16-byte chunk index is XOR-swizzled with (r & 7).
  for (int i = threadIdx.x; i < AUXT_BYTES / 2; i += THREADS) {
    int half = i / (AUX_SUB / 2), e = i % (AUX_SUB / 2), r = e / 64;
    reinterpret_cast<__half*>(smem + OFF_AUX + half * AUX_SUB)[e] = __float2half(r == 0 ? 1.f : 0.f);   // constant rows: swizzle-invariant
  }
  for (int i = threadIdx.x; i < P_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(smem + OFF_P)[i] = 0u;
  tc::fence_proxy_async();
  if (warp == 1) { tc::tmem_alloc(tmem_ptr, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    // ===================== TMA producer (lane 0) + L2 prefetchers (lanes 1..31) =====================
    // The 4-slot ring holds 64 KB in flight.  Optional experiment: the idle lanes touch the NEXT tile's plane rows
    // with prefetch.global.L2, paced one tile ahead of the producer by the __syncwarp below (see note further down).
    uint32_t it = 0;
    auto load_slot = [&](int plane, int c0, int row) {
      const int s = it % NSLOT;
      tc::mbar_wait(&empty[s], ((it / NSLOT) & 1) ^ 1);
      tc::mbar_expect_tx(&full[s], SLOT_BYTES);
      tc::tma_load_2d(smem + s * SLOT_BYTES, &tmap_x, 0, (plane * 4 + c0 / 64) * plane_rows + row, &full[s]);
      ++it;
    };
    auto job_s = [&](int i) {
      const int row = t * P + (chunk + i * chunks) * TILE_M;
      const int q0 = ps.enabled ? 0 : 2;                                                              // separable pos: S reads the x planes
      for (int ks = 0; ks < 4; ++ks) { load_slot(q0, ks * 64, row); load_slot(q0 + 1, ks * 64, row); }  // (x+pos) or x: hi, lo
    };
    auto job_z = [&](int i) {
      const int row = t * P + (chunk + i * chunks) * TILE_M;
      for (int mt = 0; mt < 2; ++mt)
        for (int pl = 0; pl < 2; ++pl) { load_slot(pl, (2 * mt) * 64, row); load_slot(pl, (2 * mt + 1) * 64, row); }   // x hi / lo, 128 channels
    };
    auto prefetch_tile = [&](int i) {            // lanes 1..31: 4 planes x 128 rows x 512 B = 2048 lines of 128 B
      const long row = (long)t * P + (long)(chunk + i * chunks) * TILE_M;
      const long max_row = (long)T * P;
      for (int ln = lane - 1; ln < 4 * TILE_M * 4; ln += 31) {
        const int pl = ln >> 9, r = (ln >> 2) & 127, seg = ln & 3;
        if (row + r < max_row) {
          const __half* ptr = planes + (((long)pl * 4 + seg) * plane_rows + row + r) * 64;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        }
      }
    };
    if (lane == 0) {
      tc::mbar_expect_tx(gfull, G_BYTES);
      for (int pl = 0; pl < 2; ++pl)
        for (int ks = 0; ks < 4; ++ks)
          tc::tma_load_2d(smem + OFF_G + (pl * 4 + ks) * G_SUB, &tmap_g, ks * 64, (t * 2 + pl) * NROW, gfull);
      job_s(0);
    }
    // Measured on B200 (1024x2048, level 3): both an L2 prefetch through the TMA engine and this LSU prefetch made the
    // kernel SLOWER (0.43 -> 0.51 ms/step): it already streams the planes at ~4.3 TB/s, the extra requests only add
    // DRAM contention.  Kept for experiments (SLOTVPS_TC_DEBUG bit 3), off by default.
    const bool do_prefetch = (dbg & 8) != 0;
    if (do_prefetch && lane != 0 && n_my > 1) prefetch_tile(1);
    __syncwarp();
    for (int i = 0; i < n_my; ++i) {
      if (lane == 0) { if (i + 1 < n_my) job_s(i + 1); if (MODE != 1) job_z(i); }
      else if (do_prefetch && i + 2 < n_my) prefetch_tile(i + 2);
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform loops, one elected lane issues (tc::elect_one) =====================
    {
      const bool el = tc::elect_one();
      tc::mbar_wait(gfull, 0);
      tc::tc_fence_after();
      const uint32_t g_base = tc::smem_u32(smem + OFF_G), p_base = tc::smem_u32(smem + OFF_P), x_base = tc::smem_u32(smem + OFF_AUX);
      uint32_t it = 0;
      // SLOTVPS_TC_DEBUG bit 4: where this thread waits (cycles, printed by CTA (0,0)).  Measured on B200 (level 3, 14 tiles per
      // CTA, ~13.3 K cycles per tile): barrier waits 14-23 % (S operands 4-7 %, Z operands 8-13 %, softmax 1-2 %, S buffer
      // 0.4 %), tcgen05.commit 8 %; the rest is MMA issue.  Hoisting every descriptor out of the loops changed nothing, issuing
      // S and Z from two threads faulted, and dropping the Z products (timing experiment) shortened the kernel by only 6 %:
      // the issuing thread and the softmax warps run at about the same per-tile rate, so both have to get faster together.
      const bool prof = (dbg & 16) != 0;
      long long wS = 0, wZ = 0, wP = 0, wE = 0, tC = 0;
      auto timed_commit = [&](uint64_t* bar) {
        if (!el) return;
        if (!prof) { tc::umma_commit(bar); return; }
        const long long t0 = clock64();
        tc::umma_commit(bar);
        tC += clock64() - t0;
      };
      const long long t_begin = clock64();
      auto timed_wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {
        if (!prof) { tc::mbar_wait(bar, parity); return; }
        const long long t0 = clock64();
        tc::mbar_wait(bar, parity);
        acc += clock64() - t0;
      };
      auto issue_s = [&](int i) {
        const int b = i & 1, u = i >> 1;
        timed_wait(&sempty[b], (u & 1) ^ 1, wE);           // softmax of tile i-2 has drained this S buffer
        tc::tc_fence_after();
        const uint32_t d = tmem_base + TM_S + b * 128;
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t dgh = tc::make_smem_desc_sw128(g_base + ks * G_SUB, 16, 1024);
          const uint64_t dgl = tc::make_smem_desc_sw128(g_base + (4 + ks) * G_SUB, 16, 1024);
          {   // (x+pos) hi against G hi and G lo
            const int s = it % NSLOT;
            timed_wait(&full[s], (it / NSLOT) & 1, wS);
            tc::tc_fence_after();
            const uint64_t da = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * SLOT_BYTES), 16, 1024);
            if (el) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                tc::umma_bf16(d, da + 2 * k, dgh + 2 * k, IDESC_S, (ks | k) != 0);
                tc::umma_bf16(d, da + 2 * k, dgl + 2 * k, IDESC_S, 1);
              }
            }
            timed_commit(&empty[s]);
            ++it;
          }
          {   // (x+pos) lo against G hi
            const int s = it % NSLOT;
            timed_wait(&full[s], (it / NSLOT) & 1, wS);
            tc::tc_fence_after();
            const uint64_t da = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * SLOT_BYTES), 16, 1024);
            if (el) {
#pragma unroll
              for (int k = 0; k < 4; ++k) tc::umma_bf16(d, da + 2 * k, dgh + 2 * k, IDESC_S, 1);
            }
            timed_commit(&empty[s]);
            ++it;
          }
        }
        timed_commit(&sfull[b]);
      };
      auto issue_z = [&](int i) {
        timed_wait(pfull, i & 1, wP);                       // P and the aux tile of tile i are in shared memory
        tc::tc_fence_after();
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t d = tmem_base + TM_Z + mt * NPAD;
          for (int pl = 0; pl < 2; ++pl) {                  // pl = 0: x hi against P hi and P lo; pl = 1: x lo against P hi
            const int s = it % NSLOT;                       // even by construction: slots (s, s+1) hold channels [128 mt, 128 mt + 128)
            timed_wait(&full[s], (it / NSLOT) & 1, wZ);
            timed_wait(&full[s + 1], ((it + 1) / NSLOT) & 1, wZ);
            tc::tc_fence_after();
            // A: x tile as MN-major operand: 64-channel groups SLOT_BYTES apart, 8-pixel groups 1024 B apart
            const uint64_t da = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * SLOT_BYTES), SLOT_BYTES, 1024);
            if (el) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {                 // 16 pixels per MMA: A advances 16 rows (2048 B), B 32 B inside its 64-pixel half
                const uint32_t poff = (k >> 2) * P_SUB + (k & 3) * 32;
                tc::umma_bf16(d, da + (uint64_t)(k * 128), tc::make_smem_desc_sw128(p_base + poff, 16, 1024), IDESC_Z, (i | pl | k) != 0);
                if (pl == 0) tc::umma_bf16(d, da + (uint64_t)(k * 128), tc::make_smem_desc_sw128(p_base + P_PLANE + poff, 16, 1024), IDESC_Z, 1);
              }
            }
            timed_commit(&empty[s]);
            timed_commit(&empty[s + 1]);
            it += 2;
          }
        }
        if (el) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {                     // aux[n, :] += (P hi + P lo)[n, px] . (1, sigma_v[px], 0...)
            const uint32_t poff = (k >> 2) * P_SUB + (k & 3) * 32;
            const uint64_t dx = tc::make_smem_desc_sw128(x_base + (k >> 2) * AUX_SUB + (k & 3) * 32, 16, 1024);
            tc::umma_bf16(tmem_base + TM_AUX, tc::make_smem_desc_sw128(p_base + poff, 16, 1024), dx, IDESC_AUX, (i | k) != 0);
            tc::umma_bf16(tmem_base + TM_AUX, tc::make_smem_desc_sw128(p_base + P_PLANE + poff, 16, 1024), dx, IDESC_AUX, 1);
          }
        }
        timed_commit(pempty);                            // P may be overwritten once these retire
      };
      issue_s(0);
      for (int i = 0; i < n_my; ++i) { if (i + 1 < n_my) issue_s(i + 1); if (MODE != 1) issue_z(i); }
      if (MODE != 1 && el) tc::umma_commit(zfull);
      if (prof && el && blockIdx.x == 0 && blockIdx.y == 0)
        printf("attn_tc tiles=%d cycles=%lld wait: S-operands %lld Z-operands %lld softmax %lld S-buffer %lld | in tcgen05.commit %lld\n", n_my, clock64() - t_begin, wS, wZ, wP, wE, tC);
    }
  } else {
    // ===================== softmax warps: two threads per pixel (slot halves) + final epilogue =====================
    // The slot-axis softmax of a pixel is thread-local apart from one max and one sum exchange between the pixel's two
    // threads.  With one thread per pixel (112 exp + 208 fp16 stores each, one warp per scheduler) this phase, not the
    // tensor pipe, set the kernel's pace (measured: 38 % tensor-pipe active); eight warps halve the per-thread chain.
    const int q = warp & 3, hf = (warp - 2) >> 2;           // TMEM lane quadrant, slot half
    const int r = q * 32 + lane;                            // pixel row of the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    constexpr int NH = NPAD - NSPLIT;                       // 64: register slots per thread (the first half uses 48 of them)
    const int nb = hf == 0 ? 0 : NSPLIT;                    // first slot of this thread
    const int nh = hf == 0 ? NSPLIT : NH;                   // slots of this thread
    uint8_t* p_half = smem + OFF_P + (r >> 6) * P_SUB;      // this pixel's 64-pixel half of P
    const int pc = (r & 63) >> 3, pe = (r & 7) * 2;         // 16-byte chunk and byte offset inside it
    __half* aux_row1 = reinterpret_cast<__half*>(smem + OFF_AUX + (r >> 6) * AUX_SUB + 128 + ((pc ^ 1) * 16) + pe);
    __half* aux_row2 = reinterpret_cast<__half*>(smem + OFF_AUX + (r >> 6) * AUX_SUB + 256 + ((pc ^ 2) * 16) + pe);
    float* xmax = xch, *xsum = xch + 256;
    for (int i = 0; i < n_my; ++i) {
      const int p = (chunk + i * chunks) * TILE_M + r;
      const bool pv = p < P;
      const float rk = pv ? __ldg(rs_k + (long)t * P + p) : 0.f;
      const float rv = pv ? __ldg(rs_v + (long)t * P + p) : 0.f;
      const int b = i & 1;
      tc::mbar_wait(&sfull[b], (i >> 1) & 1);
      tc::tc_fence_after();
      float sv[NH];
      const uint32_t ta = tmem_base + lane_addr + TM_S + b * 128 + nb;
      tc::tmem_ld32(ta, sv);
      if (hf == 0) tc::tmem_ld16(ta + 32, sv + 32);
      else tc::tmem_ld32(ta + 32, sv + 32);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&sempty[b]);                          // S buffer free for tile i+2
      if (ps.pgy != nullptr && pv) {                        // + pos . G_n from the separable tables
        const float4* gy = reinterpret_cast<const float4*>(ps.pgy + ((long)t * ps.h + p / ps.w) * NPAD + nb);
        const float4* gx = reinterpret_cast<const float4*>(ps.pgx + ((long)t * ps.w + p % ps.w) * NPAD + nb);
#pragma unroll
        for (int c = 0; c < NH / 4; ++c) {
          if (4 * c < nh) {
            const float4 a = __ldg(gy + c), bq = __ldg(gx + c);
            sv[4 * c] += a.x + bq.x; sv[4 * c + 1] += a.y + bq.y; sv[4 * c + 2] += a.z + bq.z; sv[4 * c + 3] += a.w + bq.w;
          }
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < NH; ++n) {
        const float2 c = gc[nb + (n < nh ? n : 0)];
        const float sl = (n < nh && nb + n < N) ? fmaf(sv[n] + c.x, rk, c.y) : -INFINITY;
        sv[n] = sl;
        mx = fmaxf(mx, sl);
      }
      float sum = 0.f;
      if (MODE == 2) {                                      // the maximum / denominator over ALL slot groups of this pixel
        const float2 ml = pv ? grp.ml_in[(long)t * P + p] : make_float2(0.f, 1.f);
        mx = ml.x; sum = ml.y;
#pragma unroll
        for (int n = 0; n < NH; ++n) sv[n] = __expf(sv[n] - mx);
      } else {
        xmax[hf * 128 + r] = mx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mx = fmaxf(xmax[r], xmax[128 + r]);
#pragma unroll
        for (int n = 0; n < NH; ++n) { const float e = __expf(sv[n] - mx); sv[n] = e; sum += e; }
        xsum[hf * 128 + r] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        sum = xsum[r] + xsum[128 + r];
      }
      if (MODE == 1) {
        if (pv && hf == 0) grp.ml_out[(long)t * P + p] = make_float2(mx, sum);
        continue;
      }
      const float sc = pv ? rv / sum * PSCALE : 0.f;        // A' = A * rs_v (scaled, see PSCALE)
      tc::mbar_wait(pempty, (i & 1) ^ 1);                   // Z(i-1) has finished reading P
#pragma unroll
      for (int n = 0; n < NH; ++n) {
        const int ng = nb + n;
        if (n < nh && ng < NROW) {
          const float a = pv ? sv[n] * sc : 0.f;
          const __half hi = __float2half_rn(a);
          const int off = ng * 128 + ((pc ^ (ng & 7)) * 16) + pe;
          *reinterpret_cast<__half*>(p_half + off) = hi;
          *reinterpret_cast<__half*>(p_half + P_PLANE + off) = __float2half_rn(a - __half2float(hi));
        }
      }
      if (hf == 0) {
        const float sg = pv ? 1.f / rv : 0.f;               // sigma_v, hi + lo
        const __half sh = __float2half_rn(sg);
        *aux_row1 = sh;
        *aux_row2 = __float2half_rn(sg - __half2float(sh));
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(pfull);
    }
    // ---- final epilogue: Z^T (lanes = channels; one 128-channel tile per warp group) and aux (lanes = slots) -> per-CTA partials ----
    if (MODE != 1) {
    tc::mbar_wait(zfull, 1);
    tc::tc_fence_after();
    float* Zp = Zpart + (((long)chunk * T + t) * Ntot + n0) * C;
    {
      const int mt = hf, ch = mt * 128 + r;
      for (int j = 0; j < 4; ++j) {
        float v[32];
        tc::tmem_ld32(tmem_base + lane_addr + TM_Z + mt * NPAD + j * 32, v);   // last chunk over-reads 16 columns (ignored)
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) { const int n = j * 32 + c; if (n < N) Zp[(long)n * C + ch] = v[c] * PSCALE_INV; }
      }
    }
    if (hf == 0) {
      float v[32];
      tc::tmem_ld32(tmem_base + lane_addr + TM_AUX, v);     // reads 16 columns past aux (unused)
      tc::tmem_ld_wait();
      if (r < N) {
        a1part[((long)chunk * T + t) * Ntot + n0 + r] = v[0] * PSCALE_INV;
        a0part[((long)chunk * T + t) * Ntot + n0 + r] = (v[1] + v[2]) * PSCALE_INV;
      }
    }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// Total CTAs of one launch of the main-stream tensor-core kernels (attn_tc, mask_tc).  Default: every SM, the
// single-clip latency optimum; with several clips in flight narrower grids let kernels with different bottlenecks
// (tensor pipe / HBM / LSU) of different clips run side by side on disjoint SMs (SLOTVPS_MAIN_CTAS, measured 64 > 148).
inline int main_ctas() {
  static int v = 0;
  if (v == 0) { const char* e = getenv("SLOTVPS_MAIN_CTAS"); const int x = e ? atoi(e) : 148; v = (x >= 8 && x <= 148) ? x : 148; }
  return g_prof_on ? 148 : v;                               // the per-kernel profiling pass times every kernel alone, at full width
}
namespace attn {
inline int chunks_for(int P, int T) {
  int tiles = ceil_div(P, TILE_M);
  int per = main_ctas() / (T > 0 ? T : 1);
  if (per < 1) per = 1;
  return tiles < per ? tiles : per;
}
}  // namespace attn

// combined per-pixel softmax statistics of `groups` slot groups: M = max m_g, L = sum l_g exp(m_g - M)
__global__ void __launch_bounds__(256) ml_combine_kernel(const float2* __restrict__ ml_g, long stride, int groups, long n, float2* __restrict__ out) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float M = -INFINITY;
  for (int g = 0; g < groups; ++g) M = fmaxf(M, ml_g[g * stride + i].x);
  float L = 0.f;
  for (int g = 0; g < groups; ++g) { const float2 v = ml_g[g * stride + i]; L += v.y * __expf(v.x - M); }
  out[i] = make_float2(M, L);
}

template <int MODE>
inline int attn_tc_launch(const CUtensorMap& mx, const CUtensorMap& mg, const float* g0, const float* g1, const float* rs_k, const float* rs_v,
                          float* Zpart, float* a0part, float* a1part, const __half* planes, int N, int P, int T, long rows, int chunks,
                          const PosSep& ps, const SlotGroup& grp, cudaStream_t s) {
  SV_TRY(ensure_dyn_smem((const void*)attn_tc_kernel<MODE>, attn::SMEM_BYTES));
  g_prof_grid = chunks * T;
  attn_tc_kernel<MODE><<<dim3(chunks, T), attn::THREADS, attn::SMEM_BYTES, s>>>(mx, mg, g0, g1, rs_k, rs_v, Zpart, a0part, a1part, planes, N, P, T,
                                                                                (int)rows, ceil_div(P, attn::TILE_M),
                                                                                getenv("SLOTVPS_TC_DEBUG") ? atoi(getenv("SLOTVPS_TC_DEBUG")) : 0, ps, grp);
  SV_CHECK_LAUNCH(MODE == 0 ? "attn_tc" : MODE == 1 ? "attn_tc(denominators)" : "attn_tc(group)");
  return SLOTVPS_OK;
}

// Zpart/a0part/a1part [chunks][T][N]..., returns the chunk count through *chunks_out.
// N <= 104: one pass.  N > 104: slot groups of <= 104; a denominator pass per group, the per-pixel (max, sum) of
// the groups are combined, then one accumulation pass per group writes its slots' rows of the partials.
inline int tc_attention(const TcWorkspace& ws, __half* gplanes, const float* G, const float* g0, const float* g1,
                        const float* rs_k, const float* rs_v, float* Zpart, float* a0part, float* a1part, int T, int N, int P,
                        int* chunks_out, cudaStream_t s, const PosSep& ps = PosSep(), bool gplanes_ready = false) {
  CUtensorMap mx, mg;
  const long rows = (long)T * P;
  SV_TRY(tc::make_tmap_h16_sw128(&mx, ws.planes, (uint64_t)(ps.enabled ? 2 : 4) * 4 * rows, 64, attn::TILE_M));
  const int chunks = attn::chunks_for(P, T);
  *chunks_out = chunks;
  const int groups = ceil_div(N, attn::NROW);
  const size_t gstride = (size_t)T * 2 * attn::NROW * C;
  if (groups == 1) {
    if (!gplanes_ready) {                                     // slot_pre_kernel writes the planes itself
      g_planes_kernel<<<(unsigned)(((long)T * attn::NROW * C + 255) / 256), 256, 0, s>>>(G, gplanes, N, T);
      SV_CHECK_LAUNCH("g_planes");
    }
    SV_TRY(tc::make_tmap_h16_sw128(&mg, gplanes, (uint64_t)T * 2 * attn::NROW, C, attn::NROW));
    return attn_tc_launch<0>(mx, mg, g0, g1, rs_k, rs_v, Zpart, a0part, a1part, ws.planes, N, P, T, rows, chunks, ps, SlotGroup(), s);
  }
  SV_REQUIRE(ws.ml != nullptr && !(ps.enabled && ps.tky), "slot groups need the softmax-statistics workspace and materialised (x+pos) planes");
  const int base = N / groups, extra = N % groups;
  int n0 = 0;
  for (int g = 0; g < groups; ++g) {                       // denominator passes
    const int ng = base + (g < extra ? 1 : 0);
    g_planes_kernel<<<(unsigned)(((long)T * attn::NROW * C + 255) / 256), 256, 0, s>>>(G, gplanes + g * gstride, ng, T, n0, N);
    SV_CHECK_LAUNCH("g_planes");
    SV_TRY(tc::make_tmap_h16_sw128(&mg, gplanes + g * gstride, (uint64_t)T * 2 * attn::NROW, C, attn::NROW));
    SlotGroup grp;
    grp.n0 = n0; grp.n_total = N; grp.ml_out = ws.ml + (size_t)g * ws.plane_rows;
    SV_TRY(attn_tc_launch<1>(mx, mg, g0, g1, rs_k, rs_v, Zpart, a0part, a1part, ws.planes, ng, P, T, rows, chunks, ps, grp, s));
    n0 += ng;
  }
  float2* ml_all = ws.ml + (size_t)groups * ws.plane_rows;
  ml_combine_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(ws.ml, ws.plane_rows, groups, rows, ml_all);
  SV_CHECK_LAUNCH("ml_combine");
  n0 = 0;
  for (int g = 0; g < groups; ++g) {                       // accumulation passes
    const int ng = base + (g < extra ? 1 : 0);
    SV_TRY(tc::make_tmap_h16_sw128(&mg, gplanes + g * gstride, (uint64_t)T * 2 * attn::NROW, C, attn::NROW));
    SlotGroup grp;
    grp.n0 = n0; grp.n_total = N; grp.ml_in = ml_all;
    SV_TRY(attn_tc_launch<2>(mx, mg, g0, g1, rs_k, rs_v, Zpart, a0part, a1part, ws.planes, ng, P, T, rows, chunks, ps, grp, s));
    n0 += ng;
  }
  return SLOTVPS_OK;
}

}  // namespace slotvps
