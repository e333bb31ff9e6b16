// UPSNetFPN deformable-convolution subnet: per FPN level 3 x [DeformConvWithOffset 3x3 -> GroupNorm(32) -> ReLU]
// (mmdet/models/panoptic/upsnetFPN.py:36-49, 66-70; mmdet/models/utils/deform_conv_with_offset.py; the reference op is
// deform_conv_forward_cuda, mmdet/ops/dcn/src/deform_conv_cuda.cpp:152 = bilinear im2col (deform_conv_cuda_kernel.cu:190) to an
// fp32 column buffer + cuBLAS addmm).
//
// First B200 form of this row.  Activations stay pixel-major (NHWC) between the layers, so the four bilinear corners of a tap are
// contiguous channel runs; GroupNorm + ReLU of layer l are folded into the LOADS of layer l+1 (a per-(image, channel) affine);
// the sampled columns are written once as fp16 hi/lo operand planes [2][pixels][9 C_in] (the same bytes as the reference's fp32
// column buffer) and the GEMM runs on tcgen05 through the level-fusion kernel's plain-GEMM mode (fuse_tc_kernel with y_out: TMA
// A/B stages, 3-product fp16 hi/lo, fp32 accumulation in TMEM).  The 18-channel offset convolution is a direct fp32 kernel.
// Next: gather straight into the shared-memory A stages (no column planes in HBM), offsets from a tensor-core pass.
#pragma once
#include "common.cuh"
#include "fuse_tc.cuh"

namespace slotvps {
namespace dcn {
constexpr int KT = 9;                          // 3x3 taps
constexpr int NOFF = 2 * KT;                   // offset channels
constexpr int NG = 32;                         // GroupNorm groups
constexpr float ASCALE = 16.f;                 // activations and weights both carry 2^4: their product carries fuse::WSCALE = 2^8
constexpr float GN_EPS = 1e-5f;

// [B][C][P] -> [B][P][C]
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int Cn, int P) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < Cn && p < P) ? x[((long)b * Cn + c) * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < P && c < Cn) y[((long)b * P + p) * Cn + c] = tile[tx][i];
  }
}

// offset-convolution weights [18][C][3][3] -> [tap][c][20] (18 used; rows padded for 16-byte loads)
__global__ void __launch_bounds__(256) offw_prep_kernel(const float* __restrict__ w, float* __restrict__ out, int Cn) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= KT * Cn * 20) return;
  const int o = i % 20, c = (i / 20) % Cn, tap = i / (20 * Cn);
  out[i] = o < NOFF ? w[((long)o * Cn + c) * KT + tap] : 0.f;
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

// The activation entering a layer: pixel-major rows of `ld` floats, optionally through the previous layer's GroupNorm + ReLU
// folded into a per-(image, channel) affine aff[b][2][ld] (scale, shift).
struct Act {
  const float* x; int ld; const float* aff;
};
__device__ __forceinline__ float4 act4(const Act& a, int b, long pix, int c) {
  float4 v = __ldg(reinterpret_cast<const float4*>(a.x + pix * a.ld + c));
  if (a.aff) {
    const float4 s = __ldg(reinterpret_cast<const float4*>(a.aff + (long)b * 2 * a.ld + c));
    const float4 t = __ldg(reinterpret_cast<const float4*>(a.aff + (long)b * 2 * a.ld + a.ld + c));
    v.x = fmaxf(fmaf(v.x, s.x, t.x), 0.f); v.y = fmaxf(fmaf(v.y, s.y, t.y), 0.f);
    v.z = fmaxf(fmaf(v.z, s.z, t.z), 0.f); v.w = fmaxf(fmaf(v.w, s.w, t.w), 0.f);
  }
  return v;
}

// conv_offset: regular 3x3 convolution C -> 18, padding 1 (+ bias); one thread per pixel, 32-channel weight tiles in shared memory.
// off [B][P][18]
__global__ void __launch_bounds__(256) offset_conv_kernel(const Act a, const float* __restrict__ wt /*[9][C][20]*/, const float* __restrict__ bias,
                                                          float* __restrict__ off, int Cn, int H, int W) {
  __shared__ __align__(16) float ws[KT][32][20];
  const int b = blockIdx.y, P = H * W, p = blockIdx.x * 256 + threadIdx.x;
  const bool live = p < P;
  const int y = live ? p / W : 0, x = live ? p % W : 0;
  float acc[NOFF];
#pragma unroll
  for (int o = 0; o < NOFF; ++o) acc[o] = bias[o];
  for (int c0 = 0; c0 < Cn; c0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * 32 * 20; i += 256) {
      const int o = i % 20, c = (i / 20) % 32, tap = i / 640;
      ws[tap][c][o] = wt[((long)tap * Cn + c0 + c) * 20 + o];
    }
    __syncthreads();
    if (!live) continue;
#pragma unroll 1
    for (int tap = 0; tap < KT; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const long pix = (long)b * P + (long)yy * W + xx;
#pragma unroll 2
      for (int c = 0; c < 32; c += 4) {
        const float4 v = act4(a, b, pix, c0 + c);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4* wr = reinterpret_cast<const float4*>(&ws[tap][c + e][0]);
          const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
          const float2 w4 = *reinterpret_cast<const float2*>(&ws[tap][c + e][16]);
          acc[0] = fmaf(vv[e], w0.x, acc[0]); acc[1] = fmaf(vv[e], w0.y, acc[1]); acc[2] = fmaf(vv[e], w0.z, acc[2]); acc[3] = fmaf(vv[e], w0.w, acc[3]);
          acc[4] = fmaf(vv[e], w1.x, acc[4]); acc[5] = fmaf(vv[e], w1.y, acc[5]); acc[6] = fmaf(vv[e], w1.z, acc[6]); acc[7] = fmaf(vv[e], w1.w, acc[7]);
          acc[8] = fmaf(vv[e], w2.x, acc[8]); acc[9] = fmaf(vv[e], w2.y, acc[9]); acc[10] = fmaf(vv[e], w2.z, acc[10]); acc[11] = fmaf(vv[e], w2.w, acc[11]);
          acc[12] = fmaf(vv[e], w3.x, acc[12]); acc[13] = fmaf(vv[e], w3.y, acc[13]); acc[14] = fmaf(vv[e], w3.z, acc[14]); acc[15] = fmaf(vv[e], w3.w, acc[15]);
          acc[16] = fmaf(vv[e], w4.x, acc[16]); acc[17] = fmaf(vv[e], w4.y, acc[17]);
        }
      }
    }
  }
  if (live) {
    float* dst = off + ((long)b * P + p) * NOFF;
#pragma unroll
    for (int o = 0; o < NOFF; ++o) dst[o] = acc[o];
  }
}

// bilinear im2col (deform_conv_cuda_kernel.cu:190-236 with deformable_im2col_bilinear :80-112): one item = (pixel, tap, 8 channels);
// zero outside (-1, H) x (-1, W), corners outside the map contribute zero.  planes [2][rows][9 C] fp16 hi/lo, x 2^4.
struct Off { const float* p; long bs, ps, cs; };                     // offset (b, pixel, channel) at p[b * bs + pixel * ps + channel * cs]; p == null: regular conv
__global__ void __launch_bounds__(256) dcn_im2col_kernel(const Act a, const Off off, __half* __restrict__ planes, long rows, int Cn, int H, int W, int B) {
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
        const float4 v1 = (t && lf) ? act4(a, b, base + (long)h_low * W + w_low, c) : z;
        const float4 v2 = (t && rt) ? act4(a, b, base + (long)h_low * W + w_high, c) : z;
        const float4 v3 = (bt && lf) ? act4(a, b, base + (long)h_high * W + w_low, c) : z;
        const float4 v4 = (bt && rt) ? act4(a, b, base + (long)h_high * W + w_high, c) : z;
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
// -> aff [B][2][256]: scale = gamma * rstd, shift = beta - mean * scale (channels >= Cout: 0, 0)
__global__ void __launch_bounds__(256) gn_final_kernel(const double* __restrict__ part, int slabs, int B, const float* __restrict__ gw,
                                                       const float* __restrict__ gb, float* __restrict__ aff, int P, int Cout) {
  __shared__ float mean[NG], rstd[NG];
  const int b = blockIdx.x, c = threadIdx.x, cpg = Cout / NG;
  if (c < NG) {
    double sa = 0.0, sq = 0.0;
    for (int s = 0; s < slabs; ++s) { const double* src = part + (((long)s * B + b) * NG + c) * 2; sa += src[0]; sq += src[1]; }
    const double n = (double)P * cpg, m = sa / n, var = fmax(sq / n - m * m, 0.0);
    mean[c] = (float)m; rstd[c] = (float)(1.0 / sqrt(var + (double)GN_EPS));
  }
  __syncthreads();
  float sc = 0.f, sh = 0.f;
  if (c < Cout) { sc = gw[c] * rstd[c / cpg]; sh = gb[c] - mean[c / cpg] * sc; }
  aff[(long)b * 2 * C + c] = sc; aff[(long)b * 2 * C + C + c] = sh;
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

// ---- host side ----------------------------------------------------------------------------------------------------
struct LayerPrep { float* offw; __half* wplanes; };           // [9][C_in][20], [2][256][9 C_in]
inline size_t prep_layout(const slotvps_dcn_layer* L, int n, void* base, LayerPrep* out) {
  Arena a(base, (size_t)-1);
  for (int i = 0; i < n; ++i) {
    LayerPrep lp;
    lp.offw = a.take<float>((size_t)KT * L[i].c_in * 20);
    lp.wplanes = a.take<__half>((size_t)2 * C * KT * L[i].c_in);
    if (out) out[i] = lp;
  }
  return a.off;
}
struct Ws { float *xT, *y[2], *off, *aff[2]; __half* planes; double* part; int slabs; };
inline size_t ws_layout(int cin_max, int B, int H, int W, void* base, Ws* w) {
  Arena a(base, (size_t)-1);
  const size_t rows = (size_t)B * H * W;
  Ws x;
  x.xT = a.take<float>(rows * cin_max);
  x.y[0] = a.take<float>(rows * C); x.y[1] = a.take<float>(rows * C);
  x.off = a.take<float>(rows * NOFF);
  x.aff[0] = a.take<float>((size_t)B * 2 * C); x.aff[1] = a.take<float>((size_t)B * 2 * C);
  x.planes = a.take<__half>((size_t)2 * rows * KT * cin_max + 64);
  x.slabs = ceil_div(H * W, GN_SLAB);
  x.part = a.take<double>((size_t)x.slabs * B * NG * 2);
  if (w) *w = x;
  return a.off;
}
inline int validate_layer(const slotvps_dcn_layer& l) {
  if (l.c_in <= 0 || l.c_in % 64 != 0 || l.c_in > C) return fail(SLOTVPS_EINVAL, "dcn: c_in must be a multiple of 64 and <= 256%s%s");
  if (l.c_out <= 0 || l.c_out % NG != 0 || l.c_out > C) return fail(SLOTVPS_EINVAL, "dcn: c_out must be a multiple of 32 and <= 256%s%s");
  return SLOTVPS_OK;
}
// columns + GEMM of one deformable (off != null) convolution: a -> y [rows][256] raw
inline int conv_gemm(const Act& a, const Off& off, const __half* wplanes, __half* planes, float* y, int Cn, int B, int H, int W, cudaStream_t s) {
  const long rows = (long)B * H * W, K = (long)KT * Cn;
  const long items = rows * KT * (Cn / 8);
  const int grid = (int)((items + 255) / 256 < 148L * 32 ? (items + 255) / 256 : 148L * 32);
  dcn_im2col_kernel<<<grid, 256, 0, s>>>(a, off, planes, rows, Cn, H, W, B);
  SV_CHECK_LAUNCH("dcn_im2col");
  fuse::Params prm;
  memset(&prm, 0, sizeof(prm));
  prm.rows = (int)rows; prm.P = H * W; prm.w = W; prm.h = H; prm.ksub = (int)(K / 64); prm.a_lo_row = (int)rows; prm.y_out = y;
  SV_TRY(fuse_tc_launch(planes, 2 * rows, (int)rows, (int)K, wplanes, prm, s));
  return SLOTVPS_OK;
}

}  // namespace dcn
}  // namespace slotvps
