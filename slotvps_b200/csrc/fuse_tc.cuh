// Level fusion (conv_trans over cat(up2x(prev), x), dynamic_mask_head.py:172-185) on the tensor pipe.
//
// Folded form (exact up to fp32 re-association):  W.cat(up(p), x) + b = up(Wa.p) + Wb.x + b,
// and for level 0  W.cat(x,x,x) + b = (Wa0+Wa1+Wb).x + b.  Two launches of one kernel per level:
//   "coarse"  y[p'][o]  = sum_c prev[p'][c] Wa[o][c]              (A = the previous level's fp16 planes)
//   "main"    x[p][o]   = sum_c in[p][c] Wb[o][c] + b[o] + bilinear2x(y)[p][o]
// GEMM shape: M = 128 pixels per tile (TMEM lanes), N = 256 outputs, K = 128 / 256, fp16 hi/lo x3 products,
// fp32 accumulation in TMEM (two 256-column accumulators: the epilogue of tile i overlaps the MMAs of
// tile i+1).  The "main" epilogue writes everything the rest of the path needs in one pass: the fp32 NCHW
// feature (a required output of the head) and the four fp16 operand planes x_hi, x_lo, (x+pos)_hi, (x+pos)_lo
// consumed by the attention kernels -- the 256-channel feature is never re-read to build them.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "pixel_tc.cuh"

namespace slotvps {
namespace fuse {
constexpr int TILE_M = 128;
constexpr int A_BYTES = TILE_M * 128;          // [128 px][64 ch] fp16
constexpr int B_BYTES = C * 128;               // [256 out][64 ch] fp16
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int NSTAGE = 2;
constexpr int AUX_BYTES = 4096;               // barriers, tmem pointer, bias [256], feat_bn scale / shift [2][256]
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + AUX_BYTES + 1024;
constexpr int THREADS = 608;                   // warp 0 TMA loads, warp 1 MMA, warps 2..17 epilogue (four per TMEM lane quadrant), warp 18 TMA stores
constexpr int EPI_THREADS = 512;
constexpr int STG_BYTES = TILE_M * 128;        // one staged [128 px][64 ch] fp16 sub-plane tile (16 KB); four of them live in the second stage's space
constexpr uint32_t IDESC = tc::make_idesc_f16(128, 256, 0, 0);
// conv_trans weight planes carry 2^8 (a weight of ~0.05 would have a subnormal fp16 lo plane: ~1e-6 relative instead of
// 2.4e-7); the epilogue multiplies the accumulator by 2^-8 (exact)
constexpr float WSCALE = 256.f, WSCALE_INV = 1.f / 256.f;

struct Params {
  int rows;            // T * P pixel rows of this launch
  int P, w;            // pixels per frame and width of THIS resolution
  int ksub;            // K / 64
  int a_lo_row;        // row offset of the lo plane in tmap_a (hi plane at row 0)
  int a_split;         // A operand stored as 64-channel sub-planes [2][4][a_lo_row][64] (the level planes) instead of [2][rows][K]
  const float* bias;   // [256] or null (coarse)
  // coarse output
  float* y_out;        // [rows][256] or null
  int n_out;           // plain-GEMM mode only: output columns actually computed (multiple of 16, <= 256; 0 = 256): MMA N, weight box rows
  // main outputs
  const float* y_in;   // coarse y [T * P/4][256] or null (level 0)
  float* out;          // NCHW fp32, frame t at out + t*out_bs, or null
  long out_bs;
  __half* planes;      // 4 planes with stride plane_stride rows, or null
  int x_planes_only;   // separable pos: only x_hi / x_lo are written (the kernels add pos from tables)
  int tma_store;       // main pass: operand planes leave through shared-memory staging + bulk tensor stores (one pipeline stage instead of two)
  long plane_stride;
  const float* pos;    // [256][P] per frame (pos_bs) or null
  long pos_bs;
  const float *ytab, *xtab;   // separable sine tables of this resolution or null
  const float* ytabT;         // the row table as [h][128]
  int h;
  // optional (finest level): per-pixel sum_c (bn_sc[c] x[c] + bn_sh[c])^2 as four partial sums (one per 64-channel quarter,
  // written by the four epilogue threads of a pixel) ss_out [4][rows]; the consumer adds them in a fixed order (deterministic).
  const float *bn_sc, *bn_sh;
  float* ss_out;
};
}  // namespace fuse

// input features fp32 [128][P] (one pointer per frame) -> fp16 hi/lo planes [2][T*P][128]
struct Ptr8 { const float* p[SLOTVPS_MAX_FRAMES]; };
__global__ void __launch_bounds__(256) split_in_kernel(Ptr8 src, __half* __restrict__ planes, long rows, int P) {
  __shared__ float xs[CIN][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p0 = blockIdx.x * 32, t = blockIdx.y;
  const float* __restrict__ X = src.p[t];
  const int p = p0 + lane;
  for (int c = warp; c < CIN; c += 8) xs[c][lane] = p < P ? __ldg(X + (long)c * P + p) : 0.f;
  __syncthreads();
  __half2* out = reinterpret_cast<__half2*>(planes);
#pragma unroll 4
  for (int i = 0; i < 8; ++i) {
    int idx = tid + i * 256, pp = idx >> 6, cp = idx & 63;
    if (p0 + pp >= P) continue;
    long o = ((long)t * P + p0 + pp) * (CIN / 2) + cp;
    __half h0, l0, h1, l1;
    split_bf16(xs[2 * cp][pp], h0, l0); split_bf16(xs[2 * cp + 1][pp], h1, l1);
    out[o] = __halves2half2(h0, h1);
    out[rows * (CIN / 2) + o] = __halves2half2(l0, l1);
  }
}
// Same conversion without shared memory for P % 4 == 0 and 16-byte aligned frames: a thread owns 4 consecutive
// pixels x 16 channels (16 coalesced 128-bit loads along the pixel axis) and writes, per pixel, one full 32-byte
// sector of the hi plane and one of the lo plane (256-bit stores).  Warp w of a block takes channel group w.
__global__ void __launch_bounds__(256) split_in4_kernel(Ptr8 src, __half* __restrict__ planes, long rows, int P) {
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5, t = blockIdx.y;
  const int p = (blockIdx.x * 32 + lane) * 4;
  if (p >= P) return;
  const float* __restrict__ X = src.p[t] + (long)(g * 16) * P + p;
  float4 x[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) x[c] = __ldg(reinterpret_cast<const float4*>(X + (long)c * P));
  __half* hi = planes + ((long)t * P + p) * CIN + g * 16;
  __half* lo = hi + rows * CIN;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t wh[8], wl[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float a = i == 0 ? x[2 * c].x : i == 1 ? x[2 * c].y : i == 2 ? x[2 * c].z : x[2 * c].w;
      const float b = i == 0 ? x[2 * c + 1].x : i == 1 ? x[2 * c + 1].y : i == 2 ? x[2 * c + 1].z : x[2 * c + 1].w;
      __half h0, l0, h1, l1;
      split_bf16(a, h0, l0); split_bf16(b, h1, l1);
      __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
      wh[c] = *reinterpret_cast<uint32_t*>(&hh); wl[c] = *reinterpret_cast<uint32_t*>(&ll);
    }
    tc::st_global_v8(hi + (long)i * CIN, wh);
    tc::st_global_v8(lo + (long)i * CIN, wl);
  }
}
// fp32 weight [256][ld] columns [col0, col0+K) -> fp16 hi/lo planes [2][256][K]
__global__ void __launch_bounds__(256) conv_planes_kernel(const float* __restrict__ W, int ld, int col0, int K, __half* __restrict__ out) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= C * K) return;
  int o = i / K, c = i % K;
  __half h, l;
  split_bf16(W[(long)o * ld + col0 + c] * fuse::WSCALE, h, l);       // scaled: keeps the lo plane out of fp16's subnormal range
  out[i] = h; out[(long)C * K + i] = l;
}

// MODE (compile time, so that each form keeps only its own state in registers): 0 = plain GEMM to y_out (coarse pass, DCN GEMMs),
// 1 = main pass with the warp-cooperative gather of the coarse term (w % 32 == 0), 2 = main pass, generic gather / no coarse term.
template <int MODE>
__global__ void __launch_bounds__(fuse::THREADS, 1)
fuse_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_o,
               const fuse::Params prm) {
  using namespace fuse;
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  uint8_t* aux = smem + NSTAGE * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tfull = empty + NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* sfull = reinterpret_cast<uint64_t*>(aux + 128);   // [2] staged plane pair (x hi/lo | (x+pos) hi/lo) written: 512 arrivals
  uint64_t* sempty = sfull + 2;                               // [2] the bulk store has read the pair: 1 arrival
  // With tma_store the K loop runs on ONE 96 KB stage (the tensor pipe is ~10 % busy in the main pass: the epilogue sets the pace) and
  // the second stage's space holds the four 16 KB staging tiles.
  const int nstage = prm.tma_store ? 1 : NSTAGE;
  uint8_t* stg = smem + STAGE_BYTES;
  float* bias = reinterpret_cast<float*>(aux + 256);
  float* bnv = bias + C;                                    // [2][256] feat_bn scale, shift (when prm.ss_out)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
  const int n_tiles = (prm.rows + TILE_M - 1) / TILE_M;
  if (prm.ss_out)
    for (int i = threadIdx.x; i < 2 * C; i += THREADS) bnv[i] = i < C ? prm.bn_sc[i] : prm.bn_sh[i - C];

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap_a);
    tc::tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], EPI_THREADS); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&sfull[i], EPI_THREADS); tc::mbar_init(&sempty[i], 1); }
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < C; i += THREADS) bias[i] = prm.bias ? prm.bias[i] : 0.f;
  if (warp == 1) { tc::tmem_alloc(tmem_ptr, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row = tile * TILE_M;
        for (int ks = 0; ks < prm.ksub; ++ks, ++it) {
          const int s = it % nstage;
          tc::mbar_wait(&empty[s], ((it / nstage) & 1) ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          tc::mbar_expect_tx(&full[s], prm.n_out ? 2 * A_BYTES + 2 * prm.n_out * 128 : STAGE_BYTES);
          if (prm.a_split) {
            tc::tma_load_2d(st, &tmap_a, 0, ks * prm.a_lo_row + row, &full[s]);
            tc::tma_load_2d(st + A_BYTES, &tmap_a, 0, (4 + ks) * prm.a_lo_row + row, &full[s]);
          } else {
            tc::tma_load_2d(st, &tmap_a, ks * 64, row, &full[s]);
            tc::tma_load_2d(st + A_BYTES, &tmap_a, ks * 64, prm.a_lo_row + row, &full[s]);
          }
          tc::tma_load_2d(st + 2 * A_BYTES, &tmap_w, ks * 64, 0, &full[s]);
          tc::tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmap_w, ks * 64, C, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    {                                                         // all lanes: warp-uniform issue loop, one elected lane issues (tc::elect_one)
      const bool el = tc::elect_one();
      uint32_t it = 0, ti = 0;
      const uint32_t idesc = prm.n_out ? tc::make_idesc_f16(128, prm.n_out, 0, 0) : IDESC;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const int g = ti & 1, u = ti >> 1;
        tc::mbar_wait(&tempty[g], (u & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + g * 256;
        for (int ks = 0; ks < prm.ksub; ++ks, ++it) {
          const int s = it % nstage;
          tc::mbar_wait(&full[s], (it / nstage) & 1);
          tc::tc_fence_after();
          const uint32_t a_hi = tc::smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
          const uint64_t dah = tc::make_smem_desc_sw128(a_hi, 16, 1024), dal = tc::make_smem_desc_sw128(a_lo, 16, 1024);
          const uint64_t dbh = tc::make_smem_desc_sw128(b_hi, 16, 1024), dbl = tc::make_smem_desc_sw128(b_lo, 16, 1024);
          if (el) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tc::umma_bf16(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, (ks | k) != 0);
              tc::umma_bf16(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, 1);
              tc::umma_bf16(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1);
            }
            tc::umma_commit(&empty[s]);
          }
        }
        if (el) tc::umma_commit(&tfull[g]);
      }
    }
  } else if (warp < 18) {
    // ===================== epilogue: 16 warps, four threads per pixel, 16-channel units =====================
    // The epilogue (bias, bilinear gather of the coarse term, two fp16 hi/lo splits, position add, stores) is what bounds
    // this kernel (ncu, level 3: tensor pipe 10 %, DRAM 25 %, issue slots 33 % busy with 8 epilogue warps); sixteen warps
    // working on 16-column units keep the register footprint small enough for 576 threads and double the latency hiding.
    const int q = warp & 3;
    // Round uu of a tile: ALL sixteen warps work on the 64-channel quarter uu (warp (q, jq) on its 16-channel unit jq), so a round
    // completes one [128 px][64 ch] sub-plane tile per operand plane -- the box one bulk tensor store takes from shared memory.
    const int jq = (warp - 2) >> 2;                         // 16-channel unit of this warp inside the round's 64-channel quarter
    uint32_t sn = 0;                                        // staged rounds so far (parity of the staging barriers)
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int g = ti & 1, uu_ = ti >> 1;
      const int row = tile * TILE_M + r;
      const bool rv = row < prm.rows;
      const int t = rv ? row / prm.P : 0, p = rv ? row % prm.P : 0;
      const int py = p / prm.w, px = p % prm.w;
      // bilinear x2 source taps (align_corners=False): coarse map is (h/2) x (w/2)
      int i00 = 0, i01 = 0, i10 = 0, i11 = 0;
      float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
      // cooperative gather (w % 32 == 0: a warp's 32 pixels share one image row and one frame)
      constexpr bool coop = MODE == 1;
      int cbase = 0, cy0 = 0, cy1 = 0, ccol = 0, cs0 = 0, cs1 = 0;
      float cwy0 = 0.f, cwy1 = 0.f, cwx0 = 0.f, cwx1 = 0.f;
      if (MODE != 0 && prm.y_in) {
        const int ch = prm.h / 2, cw = prm.w / 2;
        const float sy = fmaxf((py + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((px + 0.5f) * 0.5f - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = min(y0 + 1, ch - 1), x1 = min(x0 + 1, cw - 1);
        const float ly = sy - y0, lx = sx - x0;
        const int base = t * ch * cw;
        if (coop) {
          cy0 = y0; cy1 = y1; cwy1 = ly; cwy0 = 1.f - ly; cwx1 = lx; cwx0 = 1.f - lx;
          const int first = (px - lane) / 2 - 1;               // coarse column fetched by lane 0 (clamped below)
          cs0 = x0 - first; cs1 = x1 - first;
          ccol = min(max(first + lane, 0), cw - 1);
          cbase = base;
        } else {
          i00 = base + y0 * cw + x0; i01 = base + y0 * cw + x1; i10 = base + y1 * cw + x0; i11 = base + y1 * cw + x1;
          w00 = (1.f - ly) * (1.f - lx); w01 = (1.f - ly) * lx; w10 = ly * (1.f - lx); w11 = ly * lx;
        }
      }
      const float* cr0 = coop ? prm.y_in + (long)(cbase + cy0 * (prm.w / 2) + ccol) * C : nullptr;
      const float* cr1 = coop ? prm.y_in + (long)(cbase + cy1 * (prm.w / 2) + ccol) * C : nullptr;
      if (coop && rv && lane < 18) {                          // this warp's coarse-row pieces (2 rows x 4 x 64 B), requested before the accumulator wait
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(cr0 + jq * 16 + e * 64));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(cr1 + jq * 16 + e * 64));
        }
      }
      const bool staged = prm.tma_store && tile * TILE_M + TILE_M <= prm.rows;      // full tiles leave through the staging buffers
      tc::mbar_wait(&tfull[g], uu_ & 1);
      tc::tc_fence_after();
      float ss_part = 0.f;
#pragma unroll 1
      for (int uu = 0; uu < 4; ++uu) {
        const int u = uu * 4 + jq;                            // 16-channel unit: channels [16 u, 16 u + 16)
        if (prm.n_out && u * 16 >= prm.n_out) {               // columns the narrowed MMA did not compute
          if (uu == 3) { tc::tc_fence_before(); tc::mbar_arrive(&tempty[g]); }
          continue;
        }
        float v[16];
        tc::tmem_ld16(tmem_base + lane_addr + g * 256 + u * 16, v);
        tc::tmem_ld_wait();
        if (uu == 3) { tc::tc_fence_before(); tc::mbar_arrive(&tempty[g]); }      // this warp's quarter is drained
        if (!rv) continue;
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = fmaf(v[c], WSCALE_INV, bias[u * 16 + c]);
        if (MODE == 0) {
          float* dst = prm.y_out + (long)row * C + u * 16;
          tc::st_global_v8f(dst, v); tc::st_global_v8f(dst + 8, v + 8);
          continue;
        }
        if (coop) {
          // the 32 pixels of this warp lie in one image row: lanes 0..17 fetch the 18 coarse columns the row segment
          // touches (two coarse rows each, blended vertically), every lane then takes its two columns by shuffle
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float ta[8], tb[8];
            if (lane < 18) {
              tc::ld_global_nc_v8f(cr0 + u * 16 + 8 * c, ta); tc::ld_global_nc_v8f(cr1 + u * 16 + 8 * c, tb);
#pragma unroll
              for (int e = 0; e < 8; ++e) ta[e] = cwy0 * ta[e] + cwy1 * tb[e];
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) ta[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float u0 = __shfl_sync(0xffffffffu, ta[e], cs0), u1 = __shfl_sync(0xffffffffu, ta[e], cs1);
              v[8 * c + e] += cwx0 * u0 + cwx1 * u1;
            }
          }
        } else if (MODE == 2 && prm.y_in) {
          const float* a = prm.y_in + (long)i00 * C + u * 16;
          const float* b = prm.y_in + (long)i01 * C + u * 16;
          const float* cc = prm.y_in + (long)i10 * C + u * 16;
          const float* d = prm.y_in + (long)i11 * C + u * 16;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float ya[8], yb[8], yc[8], yd[8];
            tc::ld_global_nc_v8f(a + 8 * c, ya); tc::ld_global_nc_v8f(b + 8 * c, yb);
            tc::ld_global_nc_v8f(cc + 8 * c, yc); tc::ld_global_nc_v8f(d + 8 * c, yd);
            // same association as F.interpolate: (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11), weights pre-multiplied
#pragma unroll
            for (int e = 0; e < 8; ++e) v[8 * c + e] += w00 * ya[e] + w01 * yb[e] + w10 * yc[e] + w11 * yd[e];
          }
        }
        if (prm.ss_out) {
#pragma unroll
          for (int c = 0; c < 16; ++c) { const float gq = fmaf(bnv[u * 16 + c], v[c], bnv[C + u * 16 + c]); ss_part = fmaf(gq, gq, ss_part); }
          if (uu == 3) prm.ss_out[(long)jq * prm.rows + row] = ss_part;
        }
        if (prm.out) {
          float* o = prm.out + (long)t * prm.out_bs + (long)(u * 16) * prm.P + p;
#pragma unroll
          for (int c = 0; c < 16; ++c) o[(long)c * prm.P] = v[c];
        }
        if (prm.planes) {
          uint32_t hi[8], lo[8];
          auto store = [&](int plane) {
            if (staged) {
              // swizzled [128 px][64 ch] tile (row = 128 bytes, 16-byte chunk c stored at c ^ (row & 7)): what SWIZZLE_128B boxes look like
              const int pr = plane >> 1;
              tc::mbar_wait(&sempty[pr], (sn & 1) ^ 1);       // the bulk store of the previous round has read this pair's tiles
              uint8_t* d0 = stg + plane * STG_BYTES + r * 128;
              const int c0 = jq * 2, sw = r & 7;
              *reinterpret_cast<uint4*>(d0 + (((c0) ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(d0 + (((c0 + 1) ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              *reinterpret_cast<uint4*>(d0 + STG_BYTES + (((c0) ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              *reinterpret_cast<uint4*>(d0 + STG_BYTES + (((c0 + 1) ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
              tc::fence_proxy_async();
              tc::mbar_arrive(&sfull[pr]);
              return;
            }
            // sub-plane ks = u / 4 of this plane, [rows][64]: the four units of a 64-channel group complete one 128-byte row
            __half* dst = prm.planes + (((long)plane * 4 + (u >> 2)) * prm.plane_stride + row) * 64 + (u & 3) * 16;
            tc::st_global_v8(dst, hi);
            tc::st_global_v8(dst + (long)4 * prm.plane_stride * 64, lo);       // the lo plane is the next plane
          };
#pragma unroll
          for (int c = 0; c < 8; ++c) split2(v[2 * c], v[2 * c + 1], hi[c], lo[c]);
          store(0);
          if (prm.x_planes_only) { if (staged) ++sn; continue; }
          if (prm.pos) {
            const float* ps = prm.pos + (long)t * prm.pos_bs + (long)(u * 16) * prm.P + p;
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] += __ldg(ps + (long)c * prm.P);
          } else if (prm.ytab) {
            if (u < 8) {                                      // a warp's pixels share the image row: two broadcast 32-byte loads
              float tb[16];
              tc::ld_global_nc_v8f(prm.ytabT + py * 128 + u * 16, tb); tc::ld_global_nc_v8f(prm.ytabT + py * 128 + u * 16 + 8, tb + 8);
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] += tb[c];
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] += __ldg(prm.xtab + ((u - 8) * 16 + c) * prm.w + px);
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) split2(v[2 * c], v[2 * c + 1], hi[c], lo[c]);
          store(2);
          if (staged) ++sn;
        }
      }
    }
  } else if (prm.tma_store) {
    // ===================== warp 18: bulk tensor stores of the staged operand-plane tiles =====================
    const bool el = tc::elect_one();
    const int npairs = prm.x_planes_only ? 1 : 2;
    uint32_t sn = 0;
    if (el) tc::tma_prefetch_desc(&tmap_o);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (tile * TILE_M + TILE_M > prm.rows) continue;        // the partial tile is stored directly by the epilogue threads
      for (int rd = 0; rd < 4; ++rd, ++sn) {
        for (int pr = 0; pr < npairs; ++pr) {
          tc::mbar_wait(&sfull[pr], sn & 1);
          if (el) {
#pragma unroll
            for (int hl = 0; hl < 2; ++hl) {
              const int plane = pr * 2 + hl;
              tc::tma_store_2d(&tmap_o, stg + plane * STG_BYTES, 0, (int)((plane * 4 + rd) * prm.plane_stride) + tile * TILE_M);
            }
            tc::bulk_commit();
            tc::bulk_wait_read0();
            tc::mbar_arrive(&sempty[pr]);
          }
          __syncwarp();
        }
      }
    }
    if (el) tc::bulk_wait0();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

struct FuseTcWeights {
  __half *w0 = nullptr, *wa = nullptr, *wb = nullptr;   // [2][256][128], [2][256][256], [2][256][128] hi/lo planes
};
struct FuseTcWorkspace {
  __half* in_planes = nullptr;   // [2][T*Pmax][128]
  float* y = nullptr;            // [T*Pmax/4][256]
};

inline int fuse_tc_launch(const __half* a_planes, long a_rows_total, int a_lo_row, int K, const __half* w_planes, const fuse::Params& prm,
                          cudaStream_t s, int max_ctas = 148) {
  CUtensorMap ma, mw, mo;
  if (prm.a_split) SV_TRY(tc::make_tmap_h16_sw128(&ma, a_planes, (uint64_t)4 * a_rows_total, 64, fuse::TILE_M));
  else SV_TRY(tc::make_tmap_h16_sw128(&ma, a_planes, (uint64_t)a_rows_total, (uint64_t)K, fuse::TILE_M));
  SV_TRY(tc::make_tmap_h16_sw128(&mw, w_planes, (uint64_t)2 * C, (uint64_t)K, prm.n_out ? prm.n_out : C));
  if (prm.tma_store && prm.planes)
    SV_TRY(tc::make_tmap_h16_sw128(&mo, prm.planes, (uint64_t)(prm.x_planes_only ? 2 : 4) * 4 * prm.plane_stride, 64, fuse::TILE_M));
  else mo = mw;                                                      // unused
  const int mode = prm.y_out ? 0 : (prm.y_in != nullptr && (prm.w & 31) == 0) ? 1 : 2;
  const void* fn = mode == 0 ? (const void*)fuse_tc_kernel<0> : mode == 1 ? (const void*)fuse_tc_kernel<1> : (const void*)fuse_tc_kernel<2>;
  SV_TRY(ensure_dyn_smem(fn, fuse::SMEM_BYTES));
  const int n_tiles = ceil_div(prm.rows, fuse::TILE_M);
  const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
  g_prof_grid = grid;
  if (mode == 0) fuse_tc_kernel<0><<<grid, fuse::THREADS, fuse::SMEM_BYTES, s>>>(ma, mw, mo, prm);
  else if (mode == 1) fuse_tc_kernel<1><<<grid, fuse::THREADS, fuse::SMEM_BYTES, s>>>(ma, mw, mo, prm);
  else fuse_tc_kernel<2><<<grid, fuse::THREADS, fuse::SMEM_BYTES, s>>>(ma, mw, mo, prm);
  SV_CHECK_LAUNCH("fuse_tc");
  return SLOTVPS_OK;
}

}  // namespace slotvps
