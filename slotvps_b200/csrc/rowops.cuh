// Slot-side row kernels (rows = T*N slots, width C = 256): LayerNorm variants, the 8-head slot
// self-attention core, the query-axis softmax of the Video Retriever, partial-sum reduction.
// One warp owns one 256-wide row: lane l holds elements [4l,4l+4) and [128+4l, 128+4l+4).
#pragma once
#include "common.cuh"

namespace slotvps {

struct Row8 {
  float v[8];
};
__device__ __forceinline__ Row8 load_row(const float* __restrict__ p, int lane) {
  Row8 r;
  float4 a = *(const float4*)(p + lane * 4), b = *(const float4*)(p + 128 + lane * 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void store_row(float* __restrict__ p, int lane, const Row8& r) {
  *(float4*)(p + lane * 4) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *(float4*)(p + 128 + lane * 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// LayerNorm over the 256 values held by a warp (biased variance, eps inside the sqrt), two-pass
// like torch's (mean first, then centred second moment).
__device__ __forceinline__ Row8 warp_layernorm(const Row8& x, const float* __restrict__ w, const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x.v[i];
  float mu = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float d = x.v[i] - mu; q = fmaf(d, d, q); }
  float rs = rsqrtf(warp_sum(q) * (1.f / C) + LN_EPS);
  Row8 gw = load_row(w, lane), gb = load_row(b, lane), y;
#pragma unroll
  for (int i = 0; i < 8; ++i) y.v[i] = (x.v[i] - mu) * rs * gw.v[i] + gb.v[i];
  return y;
}

// y = act(LN(x (+ add)) * w[r % wmod] + b[r % wmod]) (+ post)  ; in-place allowed.
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                      const float* __restrict__ w, const float* __restrict__ b, int wmod,
                                                      const float* __restrict__ post, float* __restrict__ y, int rows, int relu) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  Row8 v = load_row(x + (long)r * C, lane);
  if (add) { Row8 a = load_row(add + (long)r * C, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] += a.v[i]; }
  int wi = r % wmod;
  Row8 o = warp_layernorm(v, w + wi * C, b + wi * C, lane);
  if (relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] = fmaxf(o.v[i], 0.f);
  }
  if (post) { Row8 a = load_row(post + (long)r * C, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) o.v[i] += a.v[i]; }
  store_row(y + (long)r * C, lane, o);
}
inline int ln_rows(const float* x, const float* add, const float* w, const float* b, int wmod, const float* post,
                   float* y, int rows, int relu, cudaStream_t s) {
  ln_rows_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, add, w, b, wmod, post, y, rows, relu);
  SV_CHECK_LAUNCH("ln_rows");
  return SLOTVPS_OK;
}

// q = LN(qraw; norm_q);  qt = q * gamma_k;  g0 = qt . bk_c;  g1 = q . beta_k
// (the slot-side half of the folded key LayerNorm, see DESIGN.md "folded attention")
__global__ void __launch_bounds__(256) q_post_kernel(const float* __restrict__ qraw, const float* __restrict__ nq_w,
                                                     const float* __restrict__ nq_b, const float* __restrict__ gamma_k,
                                                     const float* __restrict__ beta_k, const float* __restrict__ bk_c,
                                                     float* __restrict__ qt, float* __restrict__ g0, float* __restrict__ g1, int rows) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  Row8 q = warp_layernorm(load_row(qraw + (long)r * C, lane), nq_w, nq_b, lane);
  Row8 gk = load_row(gamma_k, lane), bk = load_row(beta_k, lane), bc = load_row(bk_c, lane), t;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    t.v[i] = q.v[i] * gk.v[i];
    s0 = fmaf(t.v[i], bc.v[i], s0);
    s1 = fmaf(q.v[i], bk.v[i], s1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  store_row(qt + (long)r * C, lane, t);
  if (lane == 0) { g0[r] = s0; g1[r] = s1; }
}

// o = gamma_v * (Y + bv_c * a1) + beta_v * a0 ; r = relu(LN(o; norm_o)) ; p2 = LN(p + r; norm2)
__global__ void __launch_bounds__(256) attn_post_kernel(const float* __restrict__ Y, const float* __restrict__ a0,
                                                        const float* __restrict__ a1, const float* __restrict__ p,
                                                        const float* __restrict__ gamma_v, const float* __restrict__ beta_v,
                                                        const float* __restrict__ bv_c, const float* __restrict__ no_w,
                                                        const float* __restrict__ no_b, const float* __restrict__ n2_w,
                                                        const float* __restrict__ n2_b, float* __restrict__ attn_out,
                                                        float* __restrict__ p2, int rows) {
  int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  Row8 y = load_row(Y + (long)r * C, lane), gv = load_row(gamma_v, lane), bv = load_row(beta_v, lane), bc = load_row(bv_c, lane), o;
  float s0 = a0[r], s1 = a1[r];
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = gv.v[i] * fmaf(bc.v[i], s1, y.v[i]) + bv.v[i] * s0;
  Row8 rr = warp_layernorm(o, no_w, no_b, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) rr.v[i] = fmaxf(rr.v[i], 0.f);
  if (attn_out) store_row(attn_out + (long)r * C, lane, rr);
  if (p2) {
    Row8 pp = load_row(p + (long)r * C, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) pp.v[i] += rr.v[i];
    store_row(p2 + (long)r * C, lane, warp_layernorm(pp, n2_w, n2_b, lane));
  }
}

// nn.MultiheadAttention core: qkv [R][768] -> o [R][256]; one CTA per (head, frame, quarter of the query rows); head_dim 32.
// A latency problem (25 KB of K/V per CTA, ~0.6 MFLOP): K and V come in as 16-byte loads that are all in flight at once,
// and the score / P.V loops carry four independent accumulation chains.
__global__ void __launch_bounds__(256) mha_core_kernel(const float* __restrict__ qkv, float* __restrict__ o, int n_slots, int nhead) {
  extern __shared__ float sm[];
  const int hd = blockIdx.x, t = blockIdx.y, D = 32;
  float* ks = sm;                       // [N][33]
  float* vs = ks + n_slots * 33;        // [N][33]
  float* ps = vs + n_slots * 33;        // [8 warps][N]
  const float* base = qkv + (long)t * n_slots * 3 * C;
#pragma unroll 4
  for (int i = threadIdx.x; i < n_slots * 8; i += 256) {
    const int j = i >> 3, d4 = (i & 7) * 4;
    const float4 kk = __ldg(reinterpret_cast<const float4*>(base + (long)j * 3 * C + C + hd * D + d4));
    const float4 vv = __ldg(reinterpret_cast<const float4*>(base + (long)j * 3 * C + 2 * C + hd * D + d4));
    float* kd = ks + j * 33 + d4;
    float* vd = vs + j * 33 + d4;
    kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
    vd[0] = vv.x; vd[1] = vv.y; vd[2] = vv.z; vd[3] = vv.w;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // query rows are split over blockIdx.z so that (heads x frames x gridDim.z) CTAs fill the machine
  const int i0 = blockIdx.z * 8 + warp, istep = 8 * gridDim.z;
  const float scale = rsqrtf((float)D);
  float qn = i0 < n_slots ? __ldg(base + (long)i0 * 3 * C + hd * D + lane) * scale : 0.f;     // lane = dim
  __syncthreads();
  float* pw = ps + warp * n_slots;
  for (int i = i0; i < n_slots; i += istep) {
    const float qd = qn;
    if (i + istep < n_slots) qn = __ldg(base + (long)(i + istep) * 3 * C + hd * D + lane) * scale;
    float mx = -INFINITY;
    for (int j0 = 0; j0 < n_slots; j0 += 128) {                     // four key blocks of 32 at a time: independent chains
      float sc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* kp[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) kp[b] = ks + min(j0 + 32 * b + lane, n_slots - 1) * 33;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float q = __shfl_sync(0xffffffffu, qd, d);
#pragma unroll
        for (int b = 0; b < 4; ++b) sc[b] = fmaf(q, kp[b][d], sc[b]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = j0 + 32 * b + lane;
        if (j < n_slots) { pw[j] = sc[b]; mx = fmaxf(mx, sc[b]); }
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n_slots; j += 32) { float e = expf(pw[j] - mx); pw[j] = e; sum += e; }
    sum = warp_sum(sum);
    __syncwarp();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int j = 0;
    for (; j + 4 <= n_slots; j += 4) {
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[b] = fmaf(pw[j + b], vs[(j + b) * 33 + lane], acc[b]);
    }
    for (; j < n_slots; ++j) acc[0] = fmaf(pw[j], vs[j * 33 + lane], acc[0]);
    o[((long)t * n_slots + i) * C + hd * D + lane] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) / sum;
    __syncwarp();
  }
}

// In-place softmax down each COLUMN of L [R][R] (softmax over the query axis,
// SlotsDynamicConv with softmax_dim="slots", dynamic_mask_head.py:561-562).
__global__ void __launch_bounds__(256) col_softmax_kernel(float* __restrict__ L, int R) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float mx = -INFINITY;
  if (col < R) for (int r = ry; r < R; r += 8) mx = fmaxf(mx, L[(long)r * R + col]);
  red[ry][cx] = mx;
  __syncthreads();
  mx = red[0][cx];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i][cx]);
  __syncthreads();
  float sum = 0.f;
  if (col < R) for (int r = ry; r < R; r += 8) { float e = expf(L[(long)r * R + col] - mx); L[(long)r * R + col] = e; sum += e; }
  red[ry][cx] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i][cx];
  if (col < R) for (int r = ry; r < R; r += 8) L[(long)r * R + col] /= sum;
}

// out[i] = sum_k part[k][i] in a fixed order (deterministic cross-CTA reduction of the per-CTA
// attention partial sums).
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float* __restrict__ part, float* __restrict__ out, long n, int parts) {
  long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < parts; ++k) s += part[(long)k * n + i];
  out[i] = s;
}

// Z, a0, a1 partial sums in one launch.  The `parts` partial buffers (one per attention CTA, up to 148) are summed by
// 8 threads per output element, each over a contiguous range of parts with all its loads in flight, then combined in
// group order through shared memory: a fixed association (deterministic), one L2 round trip instead of parts / 4.
__global__ void __launch_bounds__(256) reduce_attn_parts_kernel(const float* __restrict__ Zp, const float* __restrict__ a0p, const float* __restrict__ a1p,
                                                                float* __restrict__ Z, float* __restrict__ a0, float* __restrict__ a1, long nz, long na, int parts) {
  __shared__ float4 red[8][32];
  const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long nz4 = nz >> 2;                                         // nz = T * N * 256
  const long i = (long)blockIdx.x * 32 + e;                         // float4 element of Z, then scalar elements of a0 | a1
  const int per = (parts + 7) / 8, k0 = g * per, k1 = min(parts, k0 + per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < nz4) {
    const float4* src = reinterpret_cast<const float4*>(Zp) + i;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) {
      const float4 v = __ldg(src + (long)k * nz4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  } else if (i < nz4 + 2 * na) {
    long j = i - nz4;
    const float* src = j < na ? a0p : a1p;
    if (j >= na) j -= na;
#pragma unroll 4
    for (int k = k0; k < k1; ++k) s.x += __ldg(src + (long)k * na + j);
  }
  red[g][e] = s;
  __syncthreads();
  if (g != 0) return;
#pragma unroll
  for (int q = 1; q < 8; ++q) { const float4 v = red[q][e]; s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
  if (i < nz4) reinterpret_cast<float4*>(Z)[i] = s;
  else if (i < nz4 + 2 * na) {
    long j = i - nz4;
    float* dst = j < na ? a0 : a1;
    if (j >= na) j -= na;
    dst[j] = s.x;
  }
}

}  // namespace slotvps
