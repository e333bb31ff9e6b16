// Generic strided fp32 GEMM on CUDA cores with a fused epilogue.
//
// Used for the small slot-side layers (rows = T*N slots) and as the fp32 reference path of the
// pixel-side contractions (kernel_path = 1).  C[m,n] = act(alpha * sum_k A[m,k] * (B[k,n] + B2[k,n])
//                                                          + bias + up2x(U)[m,n] + R[m,n])
// All operands are addressed with element strides so every layout of the hot path
// ([out][in] weights, NCHW features, row-major slots) is reachable without copies.
#pragma once
#include "common.cuh"

namespace slotvps {

struct GemmArgs {
  const float* A = nullptr; long a_ms = 0, a_ks = 0, a_bs = 0;
  const float* B = nullptr; long b_ks = 0, b_ns = 0, b_bs = 0;
  const float* B2 = nullptr; long b2_bs = 0;          // optional addend to B (same element strides)
  float* Cm = nullptr; long c_ms = 0, c_ns = 0, c_bs = 0;
  int M = 0, N = 0, K = 0, batch = 1;
  const float* bias = nullptr; int bias_mode = 0;     // 1: bias[m], 2: bias[n]
  const float* resid = nullptr; long r_ms = 0, r_ns = 0, r_bs = 0;
  int act = 0;                                        // 0 none, 1 relu, 2 gelu(erf)
  const float* up = nullptr; int up_h = 0, up_w = 0; long up_bs = 0;  // bilinear 2x upsample-add, U is [M][up_h*up_w]
  float alpha = 1.f;
  const float* col_scale = nullptr;                   // v = (acc*alpha + bias) * col_scale[n]
  const float* affine = nullptr;                      // device [2]: v = v*affine[0] + affine[1]
};

// bilinear x2, align_corners=False (F.interpolate(scale_factor=2), dynamic_mask_head.py:178)
__device__ __forceinline__ float up2x_sample(const float* __restrict__ u, int up_h, int up_w, int n) {
  const int wf = 2 * up_w;
  const int r = n / wf, c = n - r * wf;
  float sy = fmaxf((r + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((c + 0.5f) * 0.5f - 0.5f, 0.f);
  int y0 = (int)sy, x0 = (int)sx;
  int y1 = min(y0 + 1, up_h - 1), x1 = min(x0 + 1, up_w - 1);
  float ly = sy - y0, lx = sx - x0;
  float v00 = __ldg(u + y0 * up_w + x0), v01 = __ldg(u + y0 * up_w + x1);
  float v10 = __ldg(u + y1 * up_w + x0), v11 = __ldg(u + y1 * up_w + x1);
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  constexpr int BK = 16;
  constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == 256, "256 threads");
  constexpr int PA = BM + 4, PB = BN + 4;
  __shared__ __align__(16) float As[BK][PA];
  __shared__ __align__(16) float Bs[BK][PB];
  constexpr int LA = BM * BK / 256, LB = BN * BK / 256;

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int bz = blockIdx.z;
  const float* __restrict__ A = g.A + (long)bz * g.a_bs;
  const float* __restrict__ B = g.B + (long)bz * g.b_bs;
  const float* __restrict__ B2 = g.B2 ? g.B2 + (long)bz * g.b2_bs : nullptr;
  const bool a_kfast = (g.a_ks == 1), b_nfast = (g.b_ns == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[LA], rb[LB];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * 256;
      int k, m;
      if (a_kfast) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < g.M && gk < g.K) ? __ldg(A + (long)gm * g.a_ms + (long)gk * g.a_ks) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int idx = tid + i * 256;
      int k, n;
      if (b_nfast) { n = idx % BN; k = idx / BN; } else { k = idx % BK; n = idx / BK; }
      int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < g.K) {
        long off = (long)gk * g.b_ks + (long)gn * g.b_ns;
        v = __ldg(B + off);
        if (B2) v += __ldg(B2 + off);
      }
      rb[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * 256;
      int k, m;
      if (a_kfast) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int idx = tid + i * 256;
      int k, n;
      if (b_nfast) { n = idx % BN; k = idx / BN; } else { k = idx % BK; n = idx / BK; }
      Bs[k][n] = rb[i];
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < g.K) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) *(float4*)&a[i] = *(const float4*)&As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; j += 4) *(float4*)&b[j] = *(const float4*)&Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* __restrict__ Cm = g.Cm + (long)bz * g.c_bs;
  const float* __restrict__ R = g.resid ? g.resid + (long)bz * g.r_bs : nullptr;
  const float* __restrict__ U = g.up ? g.up + (long)bz * g.up_bs : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      if (g.bias_mode == 1) v += __ldg(g.bias + m);
      else if (g.bias_mode == 2) v += __ldg(g.bias + n);
      if (g.col_scale) v *= __ldg(g.col_scale + n);
      if (g.affine) v = fmaf(v, __ldg(g.affine), __ldg(g.affine + 1));
      if (U) v += up2x_sample(U + (long)m * g.up_h * g.up_w, g.up_h, g.up_w, n);
      if (R) v += __ldg(R + (long)m * g.r_ms + (long)n * g.r_ns);
      if (g.act == 1) v = fmaxf(v, 0.f);
      else if (g.act == 2) v = gelu_erf(v);
      Cm[(long)m * g.c_ms + (long)n * g.c_ns] = v;
    }
  }
}

// Launch helper: small tiles for small problems (fills more SMs), large tiles otherwise.
inline int sgemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return SLOTVPS_OK;
  long tiles_big = (long)ceil_div(g.M, 128) * ceil_div(g.N, 128) * g.batch;
  if (tiles_big >= 148) {
    dim3 grid(ceil_div(g.N, 128), ceil_div(g.M, 128), g.batch);
    sgemm_kernel<128, 128, 8, 8><<<grid, 256, 0, s>>>(g);
  } else {
    dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 64), g.batch);
    sgemm_kernel<64, 64, 4, 4><<<grid, 256, 0, s>>>(g);
  }
  SV_CHECK_LAUNCH("sgemm");
  return SLOTVPS_OK;
}

// Y[R,O] = act(X[R,K] . W[O,K]^T + b (+ resid[R,O]))  -- an nn.Linear on row-major slots
inline int linear(const float* X, const float* W, const float* b, float* Y, int R, int K, int O, int act,
                  const float* resid, cudaStream_t s, long ldx = -1, long ldy = -1) {
  GemmArgs g;
  g.A = X; g.a_ms = ldx < 0 ? K : ldx; g.a_ks = 1;
  g.B = W; g.b_ks = 1; g.b_ns = K;
  g.Cm = Y; g.c_ms = ldy < 0 ? O : ldy; g.c_ns = 1;
  g.M = R; g.N = O; g.K = K;
  g.bias = b; g.bias_mode = b ? 2 : 0;
  g.resid = resid; g.r_ms = O; g.r_ns = 1;
  g.act = act;
  return sgemm(g, s);
}

}  // namespace slotvps

// ---------------------------------------------------------------------------------------------------
// nn.Linear on row-major slots, both operands K-contiguous ("TN"): Y[R,O] = act(X[R,K] . W[O,K]^T + b + resid)
// 32x64 tiles, BK = 32, 3-stage cp.async pipeline, float4 shared-memory reads.  Small tiles on purpose:
// R = T*N is ~200, so parallelism has to come from the output columns (and split-K for the long-K layers).
namespace slotvps {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct LinArgs {
  const float* X; long ldx;
  const float* W;            // [O][K]
  float* Y; long ldy;
  const float* bias; const float* resid; long ldr;
  int R, K, O, act;
  int ksplit;                // > 1: blockIdx.z handles K/ksplit and writes raw partials to Y + z*R*O (ldy = O)
};

__global__ void __launch_bounds__(256) linear_tn_kernel(LinArgs g) {
  constexpr int BM = 32, BN = 64, BK = 32, PK = BK + 4, ST = 3;
  __shared__ __align__(16) float As[ST][BM][PK];
  __shared__ __align__(16) float Bs[ST][BN][PK];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int klen = g.K / g.ksplit, kbeg = blockIdx.z * klen;
  const int nk = klen / BK;
  // loader mapping: one 16-byte chunk of A and two of B per thread and stage
  const int lr = tid >> 3, lc = (tid & 7) * 4;
  const float* xa = g.X + (long)min(m0 + lr, g.R - 1) * g.ldx + kbeg + lc;
  const float* wb0 = g.W + (long)min(n0 + lr, g.O - 1) * g.K + kbeg + lc;
  const float* wb1 = g.W + (long)min(n0 + lr + 32, g.O - 1) * g.K + kbeg + lc;
  auto load = [&](int st, int kt) {
    cp_async16(&As[st][lr][lc], xa + kt * BK);
    cp_async16(&Bs[st][lr][lc], wb0 + kt * BK);
    cp_async16(&Bs[st][lr + 32][lc], wb1 + kt * BK);
  };
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) { if (s < nk) load(s, s); cp_async_commit(); }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<ST - 2>();
    __syncthreads();
    if (kt + ST - 1 < nk) load((kt + ST - 1) % ST, kt + ST - 1);
    cp_async_commit();
    const int st = kt % ST;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      float4 a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = *(const float4*)&As[st][ty * 2 + i][k4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *(const float4*)&Bs[st][j * 16 + tx][k4];   // interleaved columns: conflict-free float4 reads
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
    }
  }
  float* Y = g.Y + (g.ksplit > 1 ? (long)blockIdx.z * g.R * g.O : 0);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty * 2 + i;
    if (m >= g.R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + j * 16 + tx;
      if (n >= g.O) continue;
      float v = acc[i][j];
      if (g.ksplit == 1) {
        if (g.bias) v += __ldg(g.bias + n);
        if (g.resid) v += __ldg(g.resid + (long)m * g.ldr + n);
        if (g.act == 1) v = fmaxf(v, 0.f);
        else if (g.act == 2) v = gelu_erf(v);
      }
      Y[(long)m * g.ldy + n] = v;
    }
  }
}
// sums the split-K partials in a fixed order and applies the epilogue
__global__ void __launch_bounds__(256) linear_splitk_epilogue_kernel(const float* __restrict__ part, int parts, float* __restrict__ Y, long ldy,
                                                                     const float* __restrict__ bias, const float* __restrict__ resid, long ldr,
                                                                     int R, int O, int act) {
  long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long)R * O) return;
  int m = (int)(i / O), n = (int)(i % O);
  float v = 0.f;
  for (int p = 0; p < parts; ++p) v += part[(long)p * R * O + i];
  if (bias) v += __ldg(bias + n);
  if (resid) v += __ldg(resid + (long)m * ldr + n);
  if (act == 1) v = fmaxf(v, 0.f);
  else if (act == 2) v = gelu_erf(v);
  Y[(long)m * ldy + n] = v;
}

// `scratch` (>= 4*R*O floats) enables split-K for K >= 1024
inline int linear_fast(const float* X, const float* W, const float* b, float* Y, int R, int K, int O, int act, const float* resid,
                       cudaStream_t s, long ldx = -1, long ldy = -1, float* scratch = nullptr) {
  if (ldx < 0) ldx = K;
  if (ldy < 0) ldy = O;
  const bool ok = (K % 32 == 0) && (ldx % 4 == 0) && (((uintptr_t)X | (uintptr_t)W) % 16 == 0);
  if (!ok) return linear(X, W, b, Y, R, K, O, act, resid, s, ldx, ldy);
  LinArgs g{X, ldx, W, Y, ldy, b, resid, (long)O, R, K, O, act, 1};
  dim3 grid(ceil_div(O, 64), ceil_div(R, 32), 1);
  if (scratch && K >= 1024 && (K / 4) % 32 == 0) {
    g.ksplit = 4; g.Y = scratch; g.ldy = O; grid.z = 4;
    linear_tn_kernel<<<grid, 256, 0, s>>>(g);
    SV_CHECK_LAUNCH("linear_tn(splitk)");
    linear_splitk_epilogue_kernel<<<(unsigned)(((long)R * O + 255) / 256), 256, 0, s>>>(scratch, 4, Y, ldy, b, resid, O, R, O, act);
    SV_CHECK_LAUNCH("linear_splitk_epilogue");
    return SLOTVPS_OK;
  }
  linear_tn_kernel<<<grid, 256, 0, s>>>(g);
  SV_CHECK_LAUNCH("linear_tn");
  return SLOTVPS_OK;
}

}  // namespace slotvps
