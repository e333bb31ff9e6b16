// tcgen05 / TMEM / TMA implementation of the pixel-side retriever contraction (kernel_path = 0).
//
// Precision: the <=1e-3 per-stage tolerance rules out single-pass bf16/tf32 operands,
// so every fp32 operand is split into bf16 hi + lo planes and each product is evaluated as
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (error ~2^-16, measured 1.7e-5 per stage).
//
// Data flow per level:
//   split_planes_kernel : x fp32 [T][256][P] (+pos) -> bf16 planes  xh, xl, (x+pos)h, (x+pos)l  [T*P][256]
//   stats_tc_kernel     : per 128-pixel tile, D_k = (x+pos) . Wk_c^T and D_v = x . Wv_c^T on the tensor
//                         pipe (TMA -> smem ring -> tcgen05.mma -> TMEM), epilogue reduces each pixel's
//                         256 outputs to the LayerNorm scale rs = rsqrt(mean((D+b)^2)+eps); D never
//                         leaves the SM.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace slotvps {

struct TcStageOperands {
  __nv_bfloat16* wplanes = nullptr;     // [4][256][256]: Wk_c hi, Wk_c lo, Wv_c hi, Wv_c lo (row-major [out][in])
};
struct TcWorkspace {
  __nv_bfloat16* planes = nullptr;      // [4][T*Pmax][256]: x hi, x lo, (x+pos) hi, (x+pos) lo
  float *ytab = nullptr, *xtab = nullptr;  // separable sine tables [128][hmax], [128][wmax]
  long plane_rows = 0;                  // rows allocated per plane (T*Pmax)
};

inline void tc_stage_layout(Arena& a, TcStageOperands* o) { o->wplanes = a.take<__nv_bfloat16>((size_t)4 * C * C); }
inline void tc_workspace_layout(Arena& a, const slotvps_head_desc* d, TcWorkspace* w) {
  long Pmax = 0;
  int hmax = 0, wmax = 0;
  for (int l = 0; l < d->n_levels; ++l) {
    Pmax = max(Pmax, (long)d->h[l] * d->w[l]);
    hmax = max(hmax, d->h[l]); wmax = max(wmax, d->w[l]);
  }
  w->plane_rows = (long)d->n_frames * Pmax;
  if (d->kernel_path == 0) {
    w->planes = a.take<__nv_bfloat16>((size_t)4 * w->plane_rows * C);
    w->ytab = a.take<float>((size_t)128 * hmax);
    w->xtab = a.take<float>((size_t)128 * wmax);
  }
}
// shapes the tensor-core kernels serve; everything else runs on the fp32 path
inline bool tc_supported(const slotvps_head_desc* d, int level) { return d->h[level] * d->w[level] >= 128; }

// ---- operand preparation -----------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__global__ void __launch_bounds__(256) weight_planes_kernel(const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                            __nv_bfloat16* __restrict__ out) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= C * C) return;
  __nv_bfloat16 h, l;
  split_bf16(Wk[i], h, l); out[i] = h; out[C * C + i] = l;
  split_bf16(Wv[i], h, l); out[2 * C * C + i] = h; out[3 * C * C + i] = l;
}
// separable sine tables (position_encoding.py:236-256): channels [0,128) depend on the row, [128,256) on the column
__global__ void __launch_bounds__(256) pos_tab_kernel(float* __restrict__ ytab, float* __restrict__ xtab, int h, int w) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 128 * (h + w)) return;
  const bool isy = i < 128 * h;
  int j = isy ? i : i - 128 * h;
  int n = isy ? h : w;
  int ci = j / n, r = j % n;
  float e = (float)(r + 1) / ((float)n + 1e-6f) * 6.283185307179586f;
  float a = e / powf(10000.f, (float)(2 * (ci / 2)) / 128.f);
  (isy ? ytab : xtab)[j] = (ci & 1) ? cosf(a) : sinf(a);
}
// grid (ceil(P/32), T).  pos: tensor [256][P] per frame (pos_bs stride) | tables | none
constexpr int SPLIT_SMEM = 2 * 256 * 33 * (int)sizeof(float);
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, long x_bs, const float* __restrict__ pos, long pos_bs,
                                                           const float* __restrict__ ytab, const float* __restrict__ xtab,
                                                           __nv_bfloat16* __restrict__ planes, long plane_rows, int P, int h, int w) {
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
  __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(planes);
  const long plane_stride2 = plane_rows * (C / 2);
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    int idx = tid + i * 256, pp = idx >> 7, cp = idx & 127;
    if (p0 + pp >= P) continue;
    long o = ((long)t * P + p0 + pp) * (C / 2) + cp;
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(xs[(2 * cp) * 33 + pp], h0, l0); split_bf16(xs[(2 * cp + 1) * 33 + pp], h1, l1);
    out[o] = __halves2bfloat162(h0, h1);
    out[plane_stride2 + o] = __halves2bfloat162(l0, l1);
    split_bf16(ps[(2 * cp) * 33 + pp], h0, l0); split_bf16(ps[(2 * cp + 1) * 33 + pp], h1, l1);
    out[2 * plane_stride2 + o] = __halves2bfloat162(h0, h1);
    out[3 * plane_stride2 + o] = __halves2bfloat162(l0, l1);
  }
}

// ---- LayerNorm statistics of the key / value projections on the tensor pipe ---------------------------------
namespace stats {
constexpr int TILE_M = 128;                    // pixels per tile (TMEM lanes)
constexpr int KSUB = 64;                       // channels per k-subtile = one 128-byte swizzle row
constexpr int A_BYTES = TILE_M * 128;          // 16 KB  [128 px][64 ch] bf16
constexpr int B_BYTES = C * 128;               // 32 KB  [256 out][64 ch] bf16
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // A hi, A lo, B hi, B lo = 96 KB
constexpr int NSTAGE = 2;
constexpr int AUX_BYTES = 4096;                // barriers, tmem pointer, biases
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + AUX_BYTES + 1024;   // + alignment slack
constexpr int THREADS = 192;                   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr uint32_t IDESC = tc::make_idesc_bf16(128, 256, 0, 0);
}  // namespace stats

// planes: tensor map over [4*plane_rows][256] bf16; wplanes: tensor map over [4*256][256] bf16
__global__ void __launch_bounds__(stats::THREADS, 1)
stats_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const float* __restrict__ bk_c, const float* __restrict__ bv_c, float* __restrict__ rs_k,
                float* __restrict__ rs_v, int P, int T, int plane_rows, int tiles_per_frame) {
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
          const int aq = g == 0 ? 2 : 0, bq = g == 0 ? 0 : 2;
          for (int ks = 0; ks < C / KSUB; ++ks, ++it) {
            const int s = it % NSTAGE;
            tc::mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            tc::mbar_expect_tx(&full[s], STAGE_BYTES);
            tc::tma_load_2d(st, &tmap_x, ks * KSUB, aq * plane_rows + row, &full[s]);
            tc::tma_load_2d(st + A_BYTES, &tmap_x, ks * KSUB, (aq + 1) * plane_rows + row, &full[s]);
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
        tc::mbar_arrive(&tempty[g]);
        if (p < P) (g == 0 ? rs_k : rs_v)[(long)t * P + p] = rsqrtf(ss * (1.f / C) + LN_EPS);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// ---- host side ----------------------------------------------------------------------------------------------
inline int tc_prepare_stage(const slotvps_stage_params& sp, const float* Wk_c, const float* bk_c, const float* Wv_c,
                            const float* bv_c, TcStageOperands& o, cudaStream_t s) {
  (void)sp; (void)bk_c; (void)bv_c;
  weight_planes_kernel<<<ceil_div(C * C, 256), 256, 0, s>>>(Wk_c, Wv_c, o.wplanes);
  SV_CHECK_LAUNCH("weight_planes");
  return SLOTVPS_OK;
}

// x fp32 [T][256][P] (+ pos tensor | sine tables | nothing) -> bf16 operand planes of this level
inline int tc_split_level(const float* x, long x_bs, const float* pos, long pos_bs, bool sine, const TcWorkspace& ws, int T, int h, int w,
                          cudaStream_t s) {
  const int P = h * w;
  if ((long)T * P > ws.plane_rows || !ws.planes) return fail(SLOTVPS_EWORKSPACE, "tensor-core plane workspace too small%s%s");
  static bool attr_done = false;
  if (!attr_done) {
    SV_CHECK_CUDA(cudaFuncSetAttribute(split_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPLIT_SMEM));
    attr_done = true;
  }
  if (sine && !pos) {
    pos_tab_kernel<<<ceil_div(128 * (h + w), 256), 256, 0, s>>>(ws.ytab, ws.xtab, h, w);
    SV_CHECK_LAUNCH("pos_tab");
  }
  split_planes_kernel<<<dim3(ceil_div(P, 32), T), 256, SPLIT_SMEM, s>>>(x, x_bs, pos, pos_bs, (sine && !pos) ? ws.ytab : nullptr,
                                                                         (sine && !pos) ? ws.xtab : nullptr, ws.planes, ws.plane_rows, P, h, w);
  SV_CHECK_LAUNCH("split_planes");
  return SLOTVPS_OK;
}

// rs_k, rs_v [T][P] from the planes of this level and the stage's weight planes
inline int tc_stats(const TcStageOperands& ops, const TcWorkspace& ws, const float* bk_c, const float* bv_c, float* rs_k, float* rs_v,
                    int T, int P, cudaStream_t s) {
  CUtensorMap mx, mw;
  SV_TRY(tc::make_tmap_bf16_sw128(&mx, ws.planes, (uint64_t)4 * ws.plane_rows, C, stats::TILE_M));
  SV_TRY(tc::make_tmap_bf16_sw128(&mw, ops.wplanes, (uint64_t)4 * C, C, C));
  static bool attr_done = false;
  if (!attr_done) {
    SV_CHECK_CUDA(cudaFuncSetAttribute(stats_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stats::SMEM_BYTES));
    attr_done = true;
  }
  const int tiles_per_frame = ceil_div(P, stats::TILE_M);
  const int n_tiles = T * tiles_per_frame;
  const int grid = n_tiles < 148 ? n_tiles : 148;
  stats_tc_kernel<<<grid, stats::THREADS, stats::SMEM_BYTES, s>>>(mx, mw, bk_c, bv_c, rs_k, rs_v, P, T, (int)ws.plane_rows, tiles_per_frame);
  SV_CHECK_LAUNCH("stats_tc");
  return SLOTVPS_OK;
}

}  // namespace slotvps
