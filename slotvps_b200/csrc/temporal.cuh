// Video Retriever attention over the T*N slots of a clip (dynamic_mask_head.py:550-572, SlotsDynamicConv with
// softmax_dim="slots") in fp32 on the CUDA cores -- 2 x R x R x 256 FMAs with R = T*N <= 416 is far too small and too
// oddly shaped (softmax down the QUERY axis between the two products) for the tensor pipe to matter; what matters is
// that it is two launches spread over R/8 CTAs instead of two generic GEMMs, a softmax pass and two LayerNorm passes.
//
//   tscore_kernel   A[:, u] = softmax_l(q_l . k_u) for 8 keys u per CTA
//   tav_kernel      av_l = sum_u A[l,u] v_u ;  ty = f + relu(LN(av; norm_out)) ;  y = LN(ty; norm2)      (8 rows per CTA)
#pragma once
#include "common.cuh"
#include "rowops.cuh"

namespace slotvps {
namespace temporal {
constexpr int KEYS = 8;                        // key columns per CTA
constexpr int RMAX = 1248;                     // slots per clip held in shared memory (T * N): 12 frames x 104, 4 x 312

// Both kernels are latency problems (a few hundred KB out of L2 per CTA, ~1 MFLOP): rows are taken four at a time so that
// every lane keeps eight 16-byte loads in flight, and the A.v product splits the KEY range over the warps of a CTA (each
// value row is read once per CTA, all loads independent) with a fixed-order shared-memory reduction at the end.

// sum of d[j] over the 32 lanes for 8 values in 9 shuffles (instead of 40): halves are exchanged while the set narrows;
// returns the total of value (4 b4 + 2 b3 + b2) of the calling lane (bits of the lane index), identical on its 4 lanes
__device__ __forceinline__ float reduce8(float* d, int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = h4 ? d[j] : d[j + 4], keep = h4 ? d[j + 4] : d[j];
    d[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = h3 ? d[j] : d[j + 2], keep = h3 ? d[j + 2] : d[j];
    d[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float send = h2 ? d[0] : d[1], keep = h2 ? d[1] : d[0];
    d[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  d[0] += __shfl_xor_sync(0xffffffffu, d[0], 2);
  d[0] += __shfl_xor_sync(0xffffffffu, d[0], 1);
  return d[0];
}

constexpr int TS_WARPS = 16;
// tqkv rows r*3 + {q, k, v}
__global__ void __launch_bounds__(TS_WARPS * 32) tscore_kernel(const float* __restrict__ tqkv, float* __restrict__ A, int R) {
  __shared__ float ks[KEYS][C];
  __shared__ float S[RMAX][KEYS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, u0 = blockIdx.x * KEYS;
  for (int i = threadIdx.x; i < KEYS * C; i += TS_WARPS * 32) {
    const int j = i / C, c = i % C;
    ks[j][c] = (u0 + j < R) ? tqkv[((long)(u0 + j) * 3 + 1) * C + c] : 0.f;
  }
  __syncthreads();
  const int key = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  for (int l0 = warp * 4; l0 < R; l0 += TS_WARPS * 4) {
    float4 qa[4], qb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int l = min(l0 + i, R - 1);
      qa[i] = __ldg(reinterpret_cast<const float4*>(tqkv + (long)l * 3 * C + lane * 4));
      qb[i] = __ldg(reinterpret_cast<const float4*>(tqkv + (long)l * 3 * C + 128 + lane * 4));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float d[KEYS];
#pragma unroll
      for (int j = 0; j < KEYS; ++j) {
        const float4 ka = *reinterpret_cast<const float4*>(&ks[j][lane * 4]), kb = *reinterpret_cast<const float4*>(&ks[j][128 + lane * 4]);
        d[j] = ((qa[i].x * ka.x + qa[i].y * ka.y) + (qa[i].z * ka.z + qa[i].w * ka.w)) +
               ((qb[i].x * kb.x + qb[i].y * kb.y) + (qb[i].z * kb.z + qb[i].w * kb.w));
      }
      const float tot = reduce8(d, lane);
      if ((lane & 3) == 0 && l0 + i < R) S[l0 + i][key] = tot;
    }
  }
  __syncthreads();
  // softmax down the query axis: warp j < 8 owns key u0 + j
  const int u = u0 + warp;
  if (warp < KEYS && u < R) {
    float mx = -INFINITY;
    for (int l = lane; l < R; l += 32) mx = fmaxf(mx, S[l][warp]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int l = lane; l < R; l += 32) { const float e = expf(S[l][warp] - mx); S[l][warp] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int l = lane; l < R; l += 32) A[(long)l * R + u] = S[l][warp] / sum;
  }
}

constexpr int TAV_ROWS = 8;                    // query rows per CTA = warps per CTA
constexpr int TAV_SMEM = TAV_ROWS * RMAX * 4 + TAV_ROWS * TAV_ROWS * C * 4;   // A rows + per-warp partial rows
__global__ void __launch_bounds__(TAV_ROWS * 32) tav_kernel(const float* __restrict__ A, const float* __restrict__ tqkv, const float* __restrict__ f,
                                                            const float* __restrict__ no_w, const float* __restrict__ no_b,
                                                            const float* __restrict__ n2_w, const float* __restrict__ n2_b, float* __restrict__ y, int R) {
  extern __shared__ float tav_smem[];
  float* As = tav_smem;                                             // [8 rows][RMAX]
  float* part = tav_smem + TAV_ROWS * RMAX;                         // [8 warps][8 rows][256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l0 = blockIdx.x * TAV_ROWS;
  for (int i = threadIdx.x; i < TAV_ROWS * R; i += TAV_ROWS * 32) {
    const int r = i / R, u = i % R;
    As[r * RMAX + u] = (l0 + r < R) ? A[(long)(l0 + r) * R + u] : 0.f;
  }
  __syncthreads();
  float acc[TAV_ROWS][8];
#pragma unroll
  for (int r = 0; r < TAV_ROWS; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[r][i] = 0.f;
  // warp w takes the keys u = w, w + 8, ...: each value row is loaded once per CTA
#pragma unroll 4
  for (int u = warp; u < R; u += TAV_ROWS) {
    const Row8 v = load_row(tqkv + ((long)u * 3 + 2) * C, lane);
#pragma unroll
    for (int r = 0; r < TAV_ROWS; ++r) {
      const float a = As[r * RMAX + u];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[r][i] = fmaf(a, v.v[i], acc[r][i]);
    }
  }
#pragma unroll
  for (int r = 0; r < TAV_ROWS; ++r) {
    float* dst = part + ((warp * TAV_ROWS + r) * C);
    *reinterpret_cast<float4*>(dst + lane * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    *reinterpret_cast<float4*>(dst + 128 + lane * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
  }
  __syncthreads();
  const int l = l0 + warp;                                         // warp w finishes row w: partials in warp order
  if (l >= R) return;
  Row8 av;
#pragma unroll
  for (int i = 0; i < 8; ++i) av.v[i] = 0.f;
#pragma unroll
  for (int w = 0; w < TAV_ROWS; ++w) {
    const Row8 p = load_row(part + ((w * TAV_ROWS + warp) * C), lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) av.v[i] += p.v[i];
  }
  Row8 o = warp_layernorm(av, no_w, no_b, lane);
  const Row8 fr = load_row(f + (long)l * C, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = fmaxf(o.v[i], 0.f) + fr.v[i];
  o = warp_layernorm(o, n2_w, n2_b, lane);
  store_row(y + (long)l * C, lane, o);
}

}  // namespace temporal
}  // namespace slotvps
