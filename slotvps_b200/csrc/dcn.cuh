// UPSNetFPN deformable-convolution subnet: per FPN level 3 x [DeformConvWithOffset 3x3 -> GroupNorm(32) -> ReLU]
// (mmdet/models/panoptic/upsnetFPN.py:36-49, 66-70; mmdet/models/utils/deform_conv_with_offset.py; the reference op is
// deform_conv_forward_cuda, mmdet/ops/dcn/src/deform_conv_cuda.cpp:152 = bilinear im2col (deform_conv_cuda_kernel.cu:190) to an
// fp32 column buffer + cuBLAS addmm).
//
// B200 form of this row.  Activations stay pixel-major (NHWC) between the layers, so the four bilinear corners of a tap are
// contiguous channel runs.  Per layer:
//   operand      layer 0: the NCHW -> NHWC pass also writes the fp16 hi/lo operand planes of the input; later layers: act_planes applies
//                GroupNorm + ReLU of the previous layer (a per-(image, channel) affine) ONCE -> fp32 activation + its operand planes
//   offset conv  (regular 3x3, 18 outputs) as ONE 1x1 tensor-core GEMM with N = 9 taps x 18 = 162 outputs per pixel
//                (z[p][tap][o] = W_tap[o] . a[p]; the level-fusion kernel's plain-GEMM mode) followed by a 9-tap shift-sum -- a regular
//                convolution is a sum of shifted 1x1 convolutions, so nothing is gathered and no column buffer exists for it
//   deform conv  dcn_tc_kernel: implicit GEMM -- gather warps build the bilinear-sampled A stages straight in shared memory (fp16 hi/lo,
//                128-byte swizzle), weights by TMA, 3 hi/lo products per k-step into TMEM, MMA N = c_out; NO column buffer
//                (SLOTVPS_DCN_IM2COL=1 keeps the first form for A/B runs: bilinear im2col to fp16 planes [2][pixels][9 C_in] in HBM +
//                the plain tensor-core GEMM; bit-identical results)
//   GroupNorm    statistics of the raw output: per-tile column sums / sums of squares from the implicit-GEMM epilogue (no pass over y),
//                combined in double per (image, group) in a fixed order; a slab kernel over y where a tile would straddle two images
#pragma once
#include "common.cuh"
#include "fuse_tc.cuh"
#include <stdlib.h>

namespace slotvps {
namespace dcn {
constexpr int KT = 9;                          // 3x3 taps
constexpr int NOFF = 2 * KT;                   // offset channels
constexpr int NG = 32;                         // GroupNorm groups
constexpr float ASCALE = 16.f;                 // activations and weights both carry 2^4: their product carries fuse::WSCALE = 2^8
constexpr float GN_EPS = 1e-5f;

// [B][C][P] -> [B][P][C]; with `planes` also the fp16 hi/lo operand planes [2][B P][C] x 2^4 of the same values (layer 0 of the subnet:
// saves the separate act_planes pass over the input)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int Cn, int P,
                                                           __half* __restrict__ planes = nullptr, long rows = 0) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < Cn && p < P) ? x[((long)b * Cn + c) * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < P && c < Cn) {
      const float v = tile[tx][i];
      const long o = ((long)b * P + p) * Cn + c;
      y[o] = v;
      if (planes) {
        __half h, l;
        split_bf16(v * ASCALE, h, l);
        planes[o] = h; planes[rows * Cn + o] = l;
      }
    }
  }
}

// deformable-conv weights [C_out][C][3][3] -> fp16 hi/lo planes [2][256][K = 9 C] with k = tap * C + c (rows >= C_out zero), x 2^4
__global__ void __launch_bounds__(256) dcnw_prep_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cout, int Cn) {
  const long K = (long)KT * Cn, i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long)C * K) return;
  const int o = (int)(i / K), k = (int)(i % K), tap = k / Cn, c = k % Cn;
  __half h = __float2half_rn(0.f), l = h;
  if (o < Cout) split_bf16(w[((long)o * Cn + c) * KT + tap] * ASCALE, h, l);
  out[i] = h; out[(long)C * K + i] = l;
}

// offset-convolution weights [18][C][3][3] -> fp16 hi/lo planes [2][256][C] with row n = tap * 18 + o (rows >= 162 zero), x 2^4
__global__ void __launch_bounds__(256) offw_planes_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cn) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= C * Cn) return;
  const int n = i / Cn, c = i % Cn, tap = n / NOFF, o = n % NOFF;
  __half h = __float2half_rn(0.f), l = h;
  if (n < KT * NOFF) split_bf16(w[((long)o * Cn + c) * KT + tap] * ASCALE, h, l);
  out[i] = h; out[(long)C * Cn + i] = l;
}

// The activation entering a layer, materialised once: a[row][c] = relu(src[row][c] * scale[b][c] + shift[b][c]) (aff == null: the
// values as they are) -> dst fp32 [rows][Cn] (may alias src when ld == Cn; null: not written) and fp16 hi/lo planes [2][rows][Cn] x 2^4.
// One thread per 8 channels.
__global__ void __launch_bounds__(256) act_planes_kernel(const float* src, int ld, const float* __restrict__ aff, float* dst,
                                                         __half* __restrict__ planes, long rows, int Cn, int P) {
  const int cpr = Cn / 8;
  const long n = rows * cpr;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const long row = i / cpr;
    const int c = (int)(i % cpr) * 8;
    float v[8];
    tc::ld_global_nc_v8f(src + row * ld + c, v);
    if (aff) {
      const float* sc = aff + (row / P) * 2 * C + c;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(v[e], __ldg(sc + e), __ldg(sc + C + e)), 0.f);
    }
    if (dst) tc::st_global_v8f(dst + row * Cn + c, v);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(v[2 * e] * ASCALE, v[2 * e + 1] * ASCALE, hi[e], lo[e]);
    *reinterpret_cast<uint4*>(planes + row * Cn + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(planes + (rows + row) * Cn + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// off[b][p][o] = bias[o] + sum over the 9 taps of z[b][p + (ky-1) W + (kx-1)][tap * 18 + o], neighbours outside the map skipped
// (zero padding).  One thread per (pixel, o); z rows are 256 floats.
__global__ void __launch_bounds__(256) offset_shift_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ off,
                                                           int H, int W, long rows) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * NOFF) return;
  const int o = (int)(i % NOFF);
  const long row = i / NOFF;
  const int P = H * W, p = (int)(row % P), y = p / W, x = p % W;
  float acc = __ldg(bias + o);
#pragma unroll
  for (int tap = 0; tap < KT; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    acc += __ldg(z + (row + (long)(tap / 3 - 1) * W + (tap % 3 - 1)) * C + tap * NOFF + o);
  }
  off[i] = acc;
}

// bilinear im2col (deform_conv_cuda_kernel.cu:190-236 with deformable_im2col_bilinear :80-112): one item = (pixel, tap, 8 channels);
// zero outside (-1, H) x (-1, W), corners outside the map contribute zero.  planes [2][rows][9 C] fp16 hi/lo, x 2^4.
struct Off { const float* p; long bs, ps, cs; };                     // offset (b, pixel, channel) at p[b * bs + pixel * ps + channel * cs]; p == null: regular conv
__global__ void __launch_bounds__(256) dcn_im2col_kernel(const float* __restrict__ act /*[rows][Cn]*/, const Off off, __half* __restrict__ planes, long rows, int Cn,
                                                         int H, int W, int B) {
  const int cpp = Cn / 8;                                           // 8-channel chunks per (pixel, tap)
  const long P = (long)H * W, n = (long)B * P * KT * cpp, K = (long)KT * Cn;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const int ch = (int)(i % cpp) * 8;
    const int tap = (int)((i / cpp) % KT);
    const long row = i / ((long)cpp * KT);                          // b * P + p
    const int b = (int)(row / P), p = (int)(row % P), y = p / W, x = p % W;
    float oh = 0.f, ow = 0.f;
    if (off.p) {
      const float* o = off.p + (long)b * off.bs + (long)p * off.ps + (long)(2 * tap) * off.cs;
      oh = __ldg(o); ow = __ldg(o + off.cs);
    }
    const float h_im = (float)(y - 1 + tap / 3) + oh, w_im = (float)(x - 1 + tap % 3) + ow;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
      const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im), h_high = h_low + 1, w_high = w_low + 1;
      const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
      const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
      const bool t = h_low >= 0, bt = h_high <= H - 1, lf = w_low >= 0, rt = w_high <= W - 1;
      const long base = (long)b * P;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = ch + 4 * q;
        auto ld = [&](long pix) { return __ldg(reinterpret_cast<const float4*>(act + pix * Cn + c)); };
        const float4 v1 = (t && lf) ? ld(base + (long)h_low * W + w_low) : z;
        const float4 v2 = (t && rt) ? ld(base + (long)h_low * W + w_high) : z;
        const float4 v3 = (bt && lf) ? ld(base + (long)h_high * W + w_low) : z;
        const float4 v4 = (bt && rt) ? ld(base + (long)h_high * W + w_high) : z;
        v[4 * q + 0] = w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
        v[4 * q + 1] = w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
        v[4 * q + 2] = w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
        v[4 * q + 3] = w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
      }
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(v[2 * e] * ASCALE, v[2 * e + 1] * ASCALE, hi[e], lo[e]);
    __half* dst = planes + row * K + (long)tap * Cn + ch;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + rows * K) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- implicit-GEMM deformable convolution (no column buffer) ---------------------------------------------------------------
// y[row][o] = sum_{tap, c} bilinear(act; row, tap, c) * W[o][tap * Cn + c].  One CTA works on tiles of 128 output pixels (the TMEM
// lanes); the K loop runs over (tap, 64-channel group) stages.  Sixteen GATHER warps build the A operand of a stage directly in
// shared memory: every thread owns two (pixel, 8-channel) pieces, reads the four bilinear corners as 32-byte loads (a pixel's eight
// threads cover 256 contiguous bytes per corner), blends them in fp32, splits into fp16 hi / lo and stores the 16-byte chunks in the
// 128-byte-swizzled K-major layout a TMA box would have produced.  The weight tiles of the stage arrive by TMA, one elected lane
// issues the three hi/lo products per k-step into one of two TMEM accumulators, and four epilogue warps drain the finished tile
// (x 2^-8) to y[rows][256] while the next tile is being accumulated.  The sampling geometry (floor, weights, border tests) is
// evaluated once per (pixel, tap) and reused for the Cn / 64 stages of the tap.
namespace tcg {
constexpr int TILE_M = 128;
constexpr int A_BYTES = TILE_M * 128;                 // [128 px][64 ch] fp16
constexpr int B_BYTES = C * 128;                      // up to [256 out][64 ch] fp16
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int NSTAGE = 2;
constexpr int G_WARPS = 16, E_WARPS = 4;
constexpr int THREADS = 32 * (2 + E_WARPS + G_WARPS); // 704: warp 0 TMA, warp 1 MMA, warps 2..5 epilogue, warps 6..21 gather
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 4096 + 8192 + 1024;    // stages, barriers, GroupNorm partials [4][256][2], alignment
struct Params {
  const float* act; int Cn;                           // activation [rows][Cn] fp32
  Off off;                                            // sampling offsets (p == null: regular convolution)
  float* y;                                           // [rows][256]
  int rows, P, H, W, n_out;
  float* gn_part;                                     // optional [tiles][256][2]: per-tile column sums and sums of squares of y (GroupNorm statistics)
};
}  // namespace tcg

__global__ void __launch_bounds__(tcg::THREADS, 1) dcn_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const tcg::Params prm) {
  using namespace tcg;
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  // Stage = A hi/lo (32 KB) + the weight rows actually used (n_out x 128 B per plane): with n_out <= 128 the CTA asks for 141 KB of shared
  // memory instead of 205 KB, which leaves ~90 KB instead of ~28 KB of L1 for the gather (measured: dcn_tc 2.48 -> 2.32 ms at level 0).
  const int b_bytes = prm.n_out * 128, stage_bytes = 2 * A_BYTES + 2 * b_bytes;
  uint8_t* aux = smem + NSTAGE * stage_bytes;
  uint64_t* fullb = reinterpret_cast<uint64_t*>(aux);       // [2] weight tiles landed (TMA transaction bytes)
  uint64_t* fulla = fullb + NSTAGE;                         // [2] gathered A tiles written: one arrival per gather warp
  uint64_t* empty = fulla + NSTAGE;                         // [2] the MMAs that read the stage have retired
  uint64_t* tfull = empty + NSTAGE;                         // [2] accumulator complete
  uint64_t* tempty = tfull + 2;                             // [2] accumulator drained: 128 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int n_tiles = (prm.rows + TILE_M - 1) / TILE_M;
  const int ncs = prm.Cn / 64, nks = KT * ncs;
  // K-loop order: stage ks = tap * ncs + cs (tap-major): the sampling geometry is set up once per tap.  The cs-major order (consecutive
  // stages sample the same channel slice one pixel / one row apart, so their corner rows overlap in L1) was measured and is slower
  // (5.46 vs 5.01 ms): the geometry -- an offset load and its dependent arithmetic -- then sits in front of every stage's loads.
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&fullb[i], 1); tc::mbar_init(&fulla[i], G_WARPS); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 32 * E_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc(tmem_ptr, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int ks = 0; ks < nks; ++ks, ++it) {
          const int s = it % NSTAGE;
          tc::mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
          uint8_t* st = smem + s * stage_bytes + 2 * A_BYTES;
          const int kcol = ks * 64;                                              // column of this stage in the [n][tap * Cn + c] weight planes
          tc::mbar_expect_tx(&fullb[s], 2 * b_bytes);
          tc::tma_load_2d(st, &tmap_w, kcol, 0, &fullb[s]);
          tc::tma_load_2d(st + b_bytes, &tmap_w, kcol, C, &fullb[s]);
        }
    }
  } else if (warp == 1) {
    const bool el = tc::elect_one();
    const uint32_t idesc = tc::make_idesc_f16(128, prm.n_out, 0, 0);
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int g = ti & 1, u = ti >> 1;
      tc::mbar_wait(&tempty[g], (u & 1) ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + g * 256;
      for (int ks = 0; ks < nks; ++ks, ++it) {
        const int s = it % NSTAGE;
        tc::mbar_wait(&fullb[s], (it / NSTAGE) & 1);
        tc::mbar_wait(&fulla[s], (it / NSTAGE) & 1);
        tc::tc_fence_after();
        const uint32_t a_hi = tc::smem_u32(smem + s * stage_bytes), a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + b_bytes;
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
  } else if (warp < 2 + E_WARPS) {
    // ---- epilogue: one thread per pixel row, 32 columns at a time; optionally the tile's column sums / sums of squares (the
    //      GroupNorm statistics of the raw output: no separate pass over y) ----
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float* gsm = reinterpret_cast<float*>(aux + 4096);          // [4 quadrants][256][2]
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int g = ti & 1, u = ti >> 1;
      const long row = (long)tile * TILE_M + r;
      tc::mbar_wait(&tfull[g], u & 1);
      tc::tc_fence_after();
      for (int j = 0; j < prm.n_out; j += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + lane_addr + g * 256 + j, v);
        tc::tmem_ld_wait();
        if (j + 32 >= prm.n_out) { tc::tc_fence_before(); tc::mbar_arrive(&tempty[g]); }
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] *= fuse::WSCALE_INV;  // rows past the end were gathered as zeros: they add nothing below
        if (row < prm.rows) {
          float* dst = prm.y + row * C + j;
#pragma unroll
          for (int c = 0; c < 4; ++c) tc::st_global_v8f(dst + 8 * c, v + 8 * c);
        }
        if (prm.gn_part) {
          // transpose-reduce over the warp's 32 rows: after five halving exchanges lane l holds column j + l (31 shuffles per quantity)
          float w2[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) w2[c] = v[c] * v[c];
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const float s1 = up ? v[i] : v[i + off], s2 = up ? w2[i] : w2[i + off];
              const float r1 = __shfl_xor_sync(0xffffffffu, s1, off), r2 = __shfl_xor_sync(0xffffffffu, s2, off);
              v[i] = (up ? v[i + off] : v[i]) + r1;
              w2[i] = (up ? w2[i + off] : w2[i]) + r2;
            }
          }
          gsm[(q * C + j + lane) * 2] = v[0]; gsm[(q * C + j + lane) * 2 + 1] = w2[0];
        }
      }
      if (prm.gn_part) {
        asm volatile("bar.sync 2, 128;" ::: "memory");          // the four epilogue warps
        const int t128 = (warp - 2) * 32 + lane;
        for (int c = t128; c < prm.n_out; c += 128) {
          const float a = ((gsm[c * 2] + gsm[(C + c) * 2]) + gsm[(2 * C + c) * 2]) + gsm[(3 * C + c) * 2];
          const float b = ((gsm[c * 2 + 1] + gsm[(C + c) * 2 + 1]) + gsm[(2 * C + c) * 2 + 1]) + gsm[(3 * C + c) * 2 + 1];
          *reinterpret_cast<float2*>(prm.gn_part + ((long)tile * C + c) * 2) = make_float2(a, b);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");          // gsm is free for the next tile
      }
    }
  } else {
    // ---- gather: thread -> (pixel tg / 8 [+ 64], 8-channel chunk tg % 8) ----
    const int tg = (warp - 2 - E_WARPS) * 32 + lane, chunk = tg & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int py[2], px[2], pbase[2];                           // pixel coordinates, element offset of the image's first pixel row
      bool pv[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const long row = (long)tile * TILE_M + (tg >> 3) + 64 * k;
        pv[k] = row < prm.rows;
        const int b = pv[k] ? (int)(row / prm.P) : 0, p = pv[k] ? (int)(row % prm.P) : 0;
        py[k] = p / prm.W; px[k] = p % prm.W; pbase[k] = b * prm.P;
      }
      int ci[2][4];                                         // element offsets of the four corners' channel rows (clamped inside the map)
      float cw[2][4];                                       // corner weights x 2^4, zero where the corner is outside
      int cur_tap = -1;
      for (int ks = 0; ks < nks; ++ks, ++it) {
        const int tap = ks / ncs, cs = ks % ncs;
        if (tap != cur_tap) {                               // sampling geometry of this tap: once per ncs stages
          cur_tap = tap;
  #pragma unroll
          for (int k = 0; k < 2; ++k) {
            float oh = 0.f, ow = 0.f;
            if (prm.off.p && pv[k]) {
              const float* o = prm.off.p + (long)(pbase[k] / prm.P) * prm.off.bs + (long)(py[k] * prm.W + px[k]) * prm.off.ps + (long)(2 * tap) * prm.off.cs;
              oh = __ldg(o); ow = __ldg(o + prm.off.cs);
            }
            const float h_im = (float)(py[k] - 1 + tap / 3) + oh, w_im = (float)(px[k] - 1 + tap % 3) + ow;
            const bool in = pv[k] && h_im > -1.f && w_im > -1.f && h_im < (float)prm.H && w_im < (float)prm.W;
            const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im), h_high = h_low + 1, w_high = w_low + 1;
            const float lh = h_im - h_low, lw = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw;
            const bool t = in && h_low >= 0, bt = in && h_high <= prm.H - 1, lf = w_low >= 0, rt = w_high <= prm.W - 1;
            const int hl = min(max(h_low, 0), prm.H - 1), hh_ = min(max(h_high, 0), prm.H - 1);
            const int wl = min(max(w_low, 0), prm.W - 1), wh = min(max(w_high, 0), prm.W - 1);
            ci[k][0] = (pbase[k] + hl * prm.W + wl) * prm.Cn; cw[k][0] = (t && lf) ? hh * hw * ASCALE : 0.f;
            ci[k][1] = (pbase[k] + hl * prm.W + wh) * prm.Cn; cw[k][1] = (t && rt) ? hh * lw * ASCALE : 0.f;
            ci[k][2] = (pbase[k] + hh_ * prm.W + wl) * prm.Cn; cw[k][2] = (bt && lf) ? lh * hw * ASCALE : 0.f;
            ci[k][3] = (pbase[k] + hh_ * prm.W + wh) * prm.Cn; cw[k][3] = (bt && rt) ? lh * lw * ASCALE : 0.f;
          }
        }
        {
          const int s = it % NSTAGE;
          const float* src = prm.act + cs * 64 + chunk * 8;
          float c0[2][8], c1[2][8], c2[2][8], c3[2][8];
#pragma unroll
          for (int k = 0; k < 2; ++k) {                      // all eight corner loads of the thread in flight before the stage is awaited
            tc::ld_global_nc_v8f(src + ci[k][0], c0[k]); tc::ld_global_nc_v8f(src + ci[k][1], c1[k]);
            tc::ld_global_nc_v8f(src + ci[k][2], c2[k]); tc::ld_global_nc_v8f(src + ci[k][3], c3[k]);
          }
          tc::mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
          uint8_t* a_hi = smem + s * stage_bytes;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // same association as the reference's (w1 v1 + w2 v2 + w3 v3 + w4 v4), deform_conv_cuda_kernel.cu:108 (weights pre-scaled by 2^4: exact)
              const float a = cw[k][0] * c0[k][2 * e] + cw[k][1] * c1[k][2 * e] + cw[k][2] * c2[k][2 * e] + cw[k][3] * c3[k][2 * e];
              const float b = cw[k][0] * c0[k][2 * e + 1] + cw[k][1] * c1[k][2 * e + 1] + cw[k][2] * c2[k][2 * e + 1] + cw[k][3] * c3[k][2 * e + 1];
              split2(a, b, hi[e], lo[e]);
            }
            const int r = (tg >> 3) + 64 * k;
            const int o = r * 128 + ((chunk ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(a_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(a_hi + A_BYTES + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&fulla[s]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// GroupNorm statistics of y [B][P][256] (first Cout channels): slab partials (fp32 over <= 512 pixels, then double), fixed order
constexpr int GN_SLAB = 512;
__global__ void __launch_bounds__(256) gn_partial_kernel(const float* __restrict__ y, double* __restrict__ part /*[slabs][B][32][2]*/, int P, int Cout) {
  __shared__ float s1[256], s2[256];
  const int b = blockIdx.y, slab = blockIdx.x, c = threadIdx.x;
  const int p0 = slab * GN_SLAB, p1 = min(P, p0 + GN_SLAB);
  float a = 0.f, q = 0.f;
  if (c < Cout)
    for (int p = p0; p < p1; ++p) { const float v = y[((long)b * P + p) * C + c]; a += v; q = fmaf(v, v, q); }
  s1[c] = a; s2[c] = q;
  __syncthreads();
  const int cpg = Cout / NG;
  if (c < NG) {
    double sa = 0.0, sq = 0.0;
    for (int e = 0; e < cpg; ++e) { sa += (double)s1[c * cpg + e]; sq += (double)s2[c * cpg + e]; }
    double* dst = part + (((long)slab * gridDim.y + b) * NG + c) * 2;
    dst[0] = sa; dst[1] = sq;
  }
}
// -> aff [B][2][256]: scale = gamma * rstd, shift = beta - mean * scale (channels >= Cout: 0, 0).  Thread (stripe = tid / 32, group = tid % 32)
// adds the slabs stripe, stripe + 8, ... in double; the eight stripes are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256) gn_final_kernel(const double* __restrict__ part, int slabs, int B, const float* __restrict__ gw,
                                                       const float* __restrict__ gb, float* __restrict__ aff, int P, int Cout) {
  __shared__ double ps[8][NG][2];
  __shared__ float mean[NG], rstd[NG];
  const int b = blockIdx.x, c = threadIdx.x, cpg = Cout / NG, grp = c & 31, stripe = c >> 5;
  double sa = 0.0, sq = 0.0;
  for (int s = stripe; s < slabs; s += 8) { const double* src = part + (((long)s * B + b) * NG + grp) * 2; sa += src[0]; sq += src[1]; }
  ps[stripe][grp][0] = sa; ps[stripe][grp][1] = sq;
  __syncthreads();
  if (c < NG) {
    sa = 0.0; sq = 0.0;
    for (int s = 0; s < 8; ++s) { sa += ps[s][c][0]; sq += ps[s][c][1]; }
    const double n = (double)P * cpg, m = sa / n, var = fmax(sq / n - m * m, 0.0);
    mean[c] = (float)m; rstd[c] = (float)(1.0 / sqrt(var + (double)GN_EPS));
  }
  __syncthreads();
  float sc = 0.f, sh = 0.f;
  if (c < Cout) { sc = gw[c] * rstd[c / cpg]; sh = gb[c] - mean[c / cpg] * sc; }
  aff[(long)b * 2 * C + c] = sc; aff[(long)b * 2 * C + C + c] = sh;
}
// Same from the per-tile partials of the implicit-GEMM epilogue: part [tiles][256][2] floats, tiles of 128 pixels, P % 128 == 0 (a tile lies
// in one image).  One block per (image, group): thread t adds the tiles t, t + 256, ... over the group's channels in double, then a
// fixed-order tree over the 256 threads (deterministic); the group's channels get their affine from the same block.
__global__ void __launch_bounds__(256) gn_final_tiles_kernel(const float* __restrict__ part, const float* __restrict__ gw,
                                                             const float* __restrict__ gb, float* __restrict__ aff, int P, int Cout) {
  __shared__ double ra[256], rq[256];
  const int b = blockIdx.x, grp = blockIdx.y, t0 = threadIdx.x, cpg = Cout / NG, tpi = P / 128;
  double sa = 0.0, sq = 0.0;
  for (int t = t0; t < tpi; t += 256) {
    const float* src = part + (((long)b * tpi + t) * C + grp * cpg) * 2;
    for (int e = 0; e < cpg; ++e) { sa += (double)src[2 * e]; sq += (double)src[2 * e + 1]; }
  }
  ra[t0] = sa; rq[t0] = sq;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t0 < o) { ra[t0] += ra[t0 + o]; rq[t0] += rq[t0 + o]; }
    __syncthreads();
  }
  const double n = (double)P * cpg, m = ra[0] / n, var = fmax(rq[0] / n - m * m, 0.0);
  const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)GN_EPS));
  if (t0 < cpg) {
    const int c = grp * cpg + t0;
    const float sc = gw[c] * rstd;
    aff[(long)b * 2 * C + c] = sc; aff[(long)b * 2 * C + C + c] = gb[c] - mean * sc;
  }
  if (grp == 0 && t0 >= Cout && t0 < C) { aff[(long)b * 2 * C + t0] = 0.f; aff[(long)b * 2 * C + C + t0] = 0.f; }     // channels >= Cout: (0, 0)
}
// out [B][Cout][P] = relu(y * scale + shift)  (NHWC rows of 256 -> NCHW), or the raw values when aff == null
__global__ void __launch_bounds__(256) act_to_nchw_kernel(const float* __restrict__ y, const float* __restrict__ aff, float* __restrict__ out, int P, int Cout) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    float v = 0.f;
    if (p < P && c < Cout) {
      v = y[((long)b * P + p) * C + c];
      if (aff) v = fmaxf(fmaf(v, aff[(long)b * 2 * C + c], aff[(long)b * 2 * C + C + c]), 0.f);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (c < Cout && p < P) out[((long)b * Cout + c) * P + p] = tile[tx][i];
  }
}

// SLOTVPS_DCN_IM2COL=1 keeps the first form (column planes in HBM + plain GEMM) for A/B measurements
inline bool use_im2col() { const char* e = getenv("SLOTVPS_DCN_IM2COL"); return e && atoi(e) != 0; }     // read per call: one process can compare both
// ---- host side ----------------------------------------------------------------------------------------------------
struct LayerPrep { __half *offw, *wplanes; };                 // [2][256][C_in] (162 rows used), [2][256][9 C_in]
inline size_t prep_layout(const slotvps_dcn_layer* L, int n, void* base, LayerPrep* out) {
  Arena a(base, (size_t)-1);
  for (int i = 0; i < n; ++i) {
    LayerPrep lp;
    lp.offw = a.take<__half>((size_t)2 * C * L[i].c_in);
    lp.wplanes = a.take<__half>((size_t)2 * C * KT * L[i].c_in);
    if (out) out[i] = lp;
  }
  return a.off;
}
// act: fp32 activation entering the current layer [rows][cin]; y: raw conv output [rows][256] (becomes the next activation in place
// when c_out == 256, else compacted into act); z: offset-GEMM output [rows][256]; aplanes: activation planes [2][rows][cin]
struct Ws { float *act, *y, *z, *off, *aff, *tpart; __half *aplanes, *planes; double* part; int slabs; };
inline size_t ws_layout(int cin_max, int B, int H, int W, void* base, Ws* w) {
  Arena a(base, (size_t)-1);
  const size_t rows = (size_t)B * H * W;
  Ws x;
  x.act = a.take<float>(rows * C);
  x.y = a.take<float>(rows * C);
  x.z = a.take<float>(rows * C);
  x.off = a.take<float>(rows * NOFF);
  x.aff = a.take<float>((size_t)B * 2 * C);
  x.aplanes = a.take<__half>((size_t)2 * rows * cin_max + 64);
  x.planes = a.take<__half>(use_im2col() ? (size_t)2 * rows * KT * cin_max + 64 : 64);      // column planes: first form only
  x.slabs = ceil_div(H * W, GN_SLAB);
  x.part = a.take<double>((size_t)x.slabs * B * NG * 2);
  x.tpart = a.take<float>((size_t)ceil_div((int)rows, 128) * C * 2);                       // per-tile GroupNorm partials of the implicit-GEMM epilogue
  if (w) *w = x;
  return a.off;
}
inline int validate_layer(const slotvps_dcn_layer& l) {
  if (l.c_in <= 0 || l.c_in % 64 != 0 || l.c_in > C) return fail(SLOTVPS_EINVAL, "dcn: c_in must be a multiple of 64 and <= 256%s%s");
  if (l.c_out <= 0 || l.c_out % NG != 0 || l.c_out > C) return fail(SLOTVPS_EINVAL, "dcn: c_out must be a multiple of 32 and <= 256%s%s");
  return SLOTVPS_OK;
}
inline int grid_for(long items) { const long g = (items + 255) / 256; return (int)(g < 148L * 32 ? g : 148L * 32); }
// plain tensor-core GEMM y[rows][256] (first n_out columns) = A[rows][K] . W[n][K]^T / 2^8 with A, W as fp16 hi/lo planes
inline int gemm(const __half* a_planes, long rows, int K, const __half* w_planes, float* y, int n_out, int H, int W, cudaStream_t s) {
  fuse::Params prm;
  memset(&prm, 0, sizeof(prm));
  prm.rows = (int)rows; prm.P = H * W; prm.w = W; prm.h = H; prm.ksub = K / 64; prm.a_lo_row = (int)rows; prm.y_out = y;
  prm.n_out = n_out >= C ? 0 : n_out;
  return fuse_tc_launch(a_planes, 2 * rows, (int)rows, K, w_planes, prm, s);
}
// implicit-GEMM form: act [rows][Cn] -> y [rows][256] raw (first c_out columns), no column planes
inline int conv_implicit(const float* act, const Off& off, const __half* wplanes, float* y, int Cn, int c_out, int B, int H, int W, cudaStream_t s,
                         float* gn_part = nullptr) {
  const long rows = (long)B * H * W;
  tcg::Params prm;
  prm.act = act; prm.Cn = Cn; prm.off = off; prm.y = y; prm.rows = (int)rows; prm.P = H * W; prm.H = H; prm.W = W;
  prm.n_out = (c_out + 15) / 16 * 16;
  prm.gn_part = gn_part;
  const int smem_bytes = tcg::NSTAGE * (2 * tcg::A_BYTES + 2 * prm.n_out * 128) + 4096 + 8192 + 1024;
  CUtensorMap mw;
  SV_TRY(tc::make_tmap_h16_sw128(&mw, wplanes, (uint64_t)2 * C, (uint64_t)KT * Cn, prm.n_out));
  SV_TRY(ensure_dyn_smem((const void*)dcn_tc_kernel, tcg::SMEM_BYTES));
  const int n_tiles = ceil_div((int)rows, tcg::TILE_M);
  const int grid = n_tiles < 148 ? n_tiles : 148;
  g_prof_grid = grid;
  dcn_tc_kernel<<<grid, tcg::THREADS, smem_bytes, s>>>(mw, prm);
  SV_CHECK_LAUNCH("dcn_tc");
  return SLOTVPS_OK;
}
// columns + GEMM of one deformable (off.p != null) convolution: act [rows][Cn] -> y [rows][256] raw (first c_out columns)
inline int conv_gemm(const float* act, const Off& off, const __half* wplanes, __half* planes, float* y, int Cn, int c_out, int B, int H, int W,
                     cudaStream_t s, float* gn_part = nullptr) {
  const long rows = (long)B * H * W;
  if (!use_im2col()) return conv_implicit(act, off, wplanes, y, Cn, c_out, B, H, W, s, gn_part);
  dcn_im2col_kernel<<<grid_for(rows * KT * (Cn / 8)), 256, 0, s>>>(act, off, planes, rows, Cn, H, W, B);
  SV_CHECK_LAUNCH("dcn_im2col");
  return gemm(planes, rows, KT * Cn, wplanes, y, (c_out + 15) / 16 * 16, H, W, s);
}

}  // namespace dcn
}  // namespace slotvps
