// fp32 CUDA-core implementation of the pixel-side retriever kernels (kernel_path = 1).
//
// This is the numerically conservative path: the tcgen05 kernels (pixel_tc.cuh) are validated
// against it at full size on the GPU, and it serves shapes the tensor-core kernels do not cover.
// Both use the same FOLDED formulation of MaskDynamicConv (dynamic_mask_head.py:423-461), see
// DESIGN.md "folded attention":
//   rs_k[p] = rsqrt(mean_o((Wk_c (x_p+pos_p) + bk_c)_o^2) + eps)      (LayerNorm scale of k)
//   rs_v[p] = rsqrt(mean_o((Wv_c  x_p        + bv_c)_o^2) + eps)      (LayerNorm scale of v)
//   S[n,p]  = rs_k[p] * ((x_p+pos_p) . G_n + g0_n) + g1_n             (== q_n . k_p)
//   A       = softmax over the slot axis n;  A' = A * rs_v[p]
//   Z[n,:]  = sum_p A'[n,p] x_p ;  a0[n] = sum_p A[n,p] ;  a1[n] = sum_p A'[n,p]
// with Wk_c/Wv_c the output-centred projection weights (so LayerNorm's mean never appears).
#pragma once
#include "common.cuh"

namespace slotvps {

// ---- LayerNorm scale of a 256->256 projection, per pixel --------------------------------------
// grid (ceil(P/32), T); x [T][256][P] (+ pos [T or 1][256][P]); Wc [256][256] row-major; rs [T][P]
__global__ void __launch_bounds__(256) proj_rstd_kernel(const float* __restrict__ x, long x_bs, const float* __restrict__ pos, long pos_bs,
                                                        const float* __restrict__ Wc, const float* __restrict__ bc,
                                                        float* __restrict__ rs, int P) {
  constexpr int BM = 256, BN = 32, BK = 16, TM = 8, TN = 4;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * BN, t = blockIdx.y;
  const float* __restrict__ X = x + (long)t * x_bs;
  const float* __restrict__ PZ = pos ? pos + (long)t * pos_bs : nullptr;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < C; k0 += BK) {
#pragma unroll
    for (int i = 0; i < BM * BK / 256; ++i) {          // A = Wc[o][c], c fast
      int idx = tid + i * 256, k = idx % BK, m = idx / BK;
      As[k][m] = __ldg(Wc + (long)m * C + k0 + k);
    }
#pragma unroll
    for (int i = 0; i < BN * BK / 256; ++i) {          // B = x[c][p], p fast
      int idx = tid + i * 256, n = idx % BN, k = idx / BN;
      int gn = n0 + n;
      float v = 0.f;
      if (gn < P) { long off = (long)(k0 + k) * P + gn; v = __ldg(X + off); if (PZ) v += __ldg(PZ + off); }
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      *(float4*)&a[0] = *(const float4*)&As[k][lane * TM];
      *(float4*)&a[4] = *(const float4*)&As[k][lane * TM + 4];
      *(float4*)&b[0] = *(const float4*)&Bs[k][warp * TN];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float ss[TN] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    float bb = __ldg(bc + lane * TM + i);
#pragma unroll
    for (int j = 0; j < TN; ++j) { float v = acc[i][j] + bb; ss[j] = fmaf(v, v, ss[j]); }
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    float s = warp_sum(ss[j]);
    int n = n0 + warp * TN + j;
    if (lane == 0 && n < P) rs[(long)t * P + n] = rsqrtf(s * (1.f / C) + LN_EPS);
  }
}

// ---- fused slot-axis-softmax attention over pixel tiles -----------------------------------------
// grid (chunks, T, NB); CTA (chunk, t, zb) walks pixel tiles chunk, chunk+chunks, ... of 64 px,
// computes S for ALL slots (NB blocks of 128), the slot softmax, and accumulates Z for slot block zb.
// G [T][N][256], g0/g1 [T][N], rs_k/rs_v [T][P]; outputs partial sums per chunk.
constexpr int ATT_PX = 64;
constexpr size_t att_fp32_smem_bytes() { return (size_t)(16 * (128 + 4) + 16 * (ATT_PX + 4) + ATT_PX * (128 + 4) + ATT_PX * (C + 4)) * sizeof(float); }

template <int NB>
__global__ void __launch_bounds__(256, 1) slot_attn_fp32_kernel(
    const float* __restrict__ x, long x_bs, const float* __restrict__ pos, long pos_bs,
    const float* __restrict__ G, const float* __restrict__ g0, const float* __restrict__ g1,
    const float* __restrict__ rs_k, const float* __restrict__ rs_v,
    float* __restrict__ Zpart, float* __restrict__ a0part, float* __restrict__ a1part,
    int N, int P, int T) {
  constexpr int BK = 16, PA = 128 + 4, PB = ATT_PX + 4, PC = C + 4;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                       // [BK][PA]   G tile  (k = channel, m = slot)
  float* Bs = As + BK * PA;               // [BK][PB]   x+pos tile (k = channel, n = pixel)
  float* Ap = Bs + BK * PB;               // [ATT_PX][PA]  A' (k = pixel, m = slot of block zb)
  float* Xs = Ap + ATT_PX * PA;           // [ATT_PX][PC]  x tile transposed (k = pixel, n = channel)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x, chunks = gridDim.x, t = blockIdx.y, zb = blockIdx.z;
  const float* __restrict__ X = x + (long)t * x_bs;
  const float* __restrict__ PZ = pos ? pos + (long)t * pos_bs : nullptr;
  const float* __restrict__ Gt = G + (long)t * N * C;
  const int tiles = ceil_div(P, ATT_PX);
  // Z micro-tile: ty = slot group of 8, tx = channel group of 16
  const int tx = tid & 15, ty = tid >> 4;
  float zacc[8][16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) zacc[i][j] = 0.f;
  float a0acc[4] = {0.f, 0.f, 0.f, 0.f}, a1acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tile = chunk; tile < tiles; tile += chunks) {
    const int p0 = tile * ATT_PX;
    // ---------------- S = G . (x+pos) for all slot blocks -------------------------------------
    float sacc[NB][4][8];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) sacc[b][i][j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += BK) {
#pragma unroll
      for (int i = 0; i < ATT_PX * BK / 256; ++i) {
        int idx = tid + i * 256, n = idx % ATT_PX, k = idx / ATT_PX;
        int gp = p0 + n;
        float v = 0.f;
        if (gp < P) { long off = (long)(k0 + k) * P + gp; v = __ldg(X + off); if (PZ) v += __ldg(PZ + off); }
        Bs[k * PB + n] = v;
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        __syncthreads();                               // previous consumers of As done (and Bs visible after 2nd sync)
#pragma unroll
        for (int i = 0; i < 128 * BK / 256; ++i) {
          int idx = tid + i * 256, k = idx % BK, m = idx / BK;
          int n = b * 128 + m;
          As[k * PA + m] = n < N ? __ldg(Gt + (long)n * C + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float a[4], bb[8];
          *(float4*)&a[0] = *(const float4*)&As[k * PA + lane * 4];
          *(float4*)&bb[0] = *(const float4*)&Bs[k * PB + warp * 8];
          *(float4*)&bb[4] = *(const float4*)&Bs[k * PB + warp * 8 + 4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) sacc[b][i][j] = fmaf(a[i], bb[j], sacc[b][i][j]);
        }
      }
      __syncthreads();
    }
    // ---------------- slot-axis softmax (thread: 4*NB slots x 8 pixels; warp spans all slots) ---
    float rk[8], rv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gp = p0 + warp * 8 + j;
      rk[j] = gp < P ? __ldg(rs_k + (long)t * P + gp) : 0.f;
      rv[j] = gp < P ? __ldg(rs_v + (long)t * P + gp) : 0.f;
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = b * 128 + lane * 4 + i;
        const bool nv = n < N;
        const float c0 = nv ? __ldg(g0 + (long)t * N + n) : 0.f, c1 = nv ? __ldg(g1 + (long)t * N + n) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sacc[b][i][j] = nv ? fmaf(sacc[b][i][j] + c0, rk[j], c1) : -INFINITY;
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool pv = (p0 + warp * 8 + j) < P;
      float mx = -INFINITY;
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 4; ++i) mx = fmaxf(mx, sacc[b][i][j]);
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int i = 0; i < 4; ++i) { float e = expf(sacc[b][i][j] - mx); sacc[b][i][j] = e; sum += e; }
      sum = warp_sum(sum);
      const float inv = pv ? 1.f / sum : 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (b != zb) continue;
        float ap[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = sacc[b][i][j] * inv;
          ap[i] = a * rv[j];
          a0acc[i] += a; a1acc[i] += ap[i];
        }
        *(float4*)&Ap[(warp * 8 + j) * PA + lane * 4] = make_float4(ap[0], ap[1], ap[2], ap[3]);
      }
    }
    // ---------------- x tile transposed to [pixel][channel] -------------------------------------
#pragma unroll 4
    for (int i = 0; i < ATT_PX * C / 256; ++i) {
      int idx = tid + i * 256, n = idx % ATT_PX, c = idx / ATT_PX;
      int gp = p0 + n;
      Xs[n * PC + c] = gp < P ? __ldg(X + (long)c * P + gp) : 0.f;
    }
    __syncthreads();
    // ---------------- Z[slot][ch] += A'^T . x -------------------------------------------------
#pragma unroll 4
    for (int k = 0; k < ATT_PX; ++k) {
      float a[8], bb[16];
      *(float4*)&a[0] = *(const float4*)&Ap[k * PA + ty * 8];
      *(float4*)&a[4] = *(const float4*)&Ap[k * PA + ty * 8 + 4];
#pragma unroll
      for (int j = 0; j < 16; j += 4) *(float4*)&bb[j] = *(const float4*)&Xs[k * PC + tx * 16 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) zacc[i][j] = fmaf(a[i], bb[j], zacc[i][j]);
    }
    __syncthreads();
  }
  // ---------------- write this CTA's partial sums ------------------------------------------------
  float* __restrict__ Zp = Zpart + ((long)chunk * T + t) * N * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = zb * 128 + ty * 8 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *(float4*)(Zp + (long)n * C + tx * 16 + j) = make_float4(zacc[i][j], zacc[i][j + 1], zacc[i][j + 2], zacc[i][j + 3]);
  }
  // a0/a1: reduce the 8 warps (pixel groups) through shared memory, fixed order
  float* red = smem;                                     // [2][8][128]
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) { red[warp * 128 + lane * 4 + i] = a0acc[i]; red[1024 + warp * 128 + lane * 4 + i] = a1acc[i]; }
  __syncthreads();
  if (tid < 128) {
    int n = zb * 128 + tid;
    if (n < N) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { s0 += red[w * 128 + tid]; s1 += red[1024 + w * 128 + tid]; }
      a0part[((long)chunk * T + t) * N + n] = s0;
      a1part[((long)chunk * T + t) * N + n] = s1;
    }
  }
}

// ---- sine position embedding (position_encoding.py:236-256) ------------------------------------
__global__ void __launch_bounds__(256) sine_pos_kernel(float* __restrict__ out, int h, int w) {
  long i = (long)blockIdx.x * 256 + threadIdx.x;
  long P = (long)h * w;
  if (i >= 256 * P) return;
  int c = (int)(i / P), p = (int)(i % P), r = p / w, col = p % w;
  int ci = c & 127;
  // fp32 arithmetic in the reference's order: embed / (last + eps) * scale, then / dim_t
  float e = (c < 128) ? (float)(r + 1) / ((float)h + 1e-6f) : (float)(col + 1) / ((float)w + 1e-6f);
  e = e * 6.283185307179586f;
  float dim_t = powf(10000.f, (float)(2 * (ci / 2)) / 128.f);
  float a = e / dim_t;
  out[i] = (ci & 1) ? cosf(a) : sinf(a);
}

// ---- per-pixel inverse L2 norm of BatchNorm'ed features (vps_temporal_slots.py:146-147) ---------
// rn[p] = 1 / max(||s*x_p + tt||_2, 1e-12);  x [256][P];  sc/sh [256] folded BN scale/shift
__global__ void __launch_bounds__(256) feat_rnorm_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                                         const float* __restrict__ sh, float* __restrict__ rn, int P) {
  int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
#pragma unroll 8
  for (int c = 0; c < C; ++c) { float g = fmaf(__ldg(sc + c), __ldg(x + (long)c * P + p), __ldg(sh + c)); s = fmaf(g, g, s); }
  rn[p] = 1.f / fmaxf(sqrtf(s), 1e-12f);
}

// Same for P % 4 == 0 and a 16-byte aligned x: 4 pixels per thread (128-bit loads along the pixel axis), the 256
// channels split over 4 thread groups of a block and combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256) feat_rnorm4_kernel(const float* __restrict__ x, const float* __restrict__ sc,
                                                          const float* __restrict__ sh, float* __restrict__ rn, int P) {
  __shared__ float4 part[4][64];
  const int q = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int p = (blockIdx.x * 64 + q) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p < P) {
#pragma unroll 8
    for (int c = g * 64; c < g * 64 + 64; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long)c * P + p));
      const float a = __ldg(sc + c), b = __ldg(sh + c);
      float t;
      t = fmaf(a, v.x, b); s.x = fmaf(t, t, s.x);
      t = fmaf(a, v.y, b); s.y = fmaf(t, t, s.y);
      t = fmaf(a, v.z, b); s.z = fmaf(t, t, s.z);
      t = fmaf(a, v.w, b); s.w = fmaf(t, t, s.w);
    }
  }
  part[g][q] = s;
  __syncthreads();
  if (g == 0 && p < P) {
    float4 t = part[0][q];
#pragma unroll
    for (int k = 1; k < 4; ++k) { const float4 u = part[k][q]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    *reinterpret_cast<float4*>(rn + p) = make_float4(1.f / fmaxf(sqrtf(t.x), 1e-12f), 1.f / fmaxf(sqrtf(t.y), 1e-12f),
                                                     1.f / fmaxf(sqrtf(t.z), 1e-12f), 1.f / fmaxf(sqrtf(t.w), 1e-12f));
  }
}

}  // namespace slotvps
