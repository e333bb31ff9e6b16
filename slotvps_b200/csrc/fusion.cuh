// Panoptic fusion on device: PostProcessPanopticInstances.forward (vps_temporal_slots.py:659-807,
// mask_removal :564-657) + the inline relabel of simple_test (:411-435), with no host round trip.
//
// Exact two-pass form of mask_removal: the per-pixel softmax over the K
// kept slots sums to 1, so at most two slots can reach pixel_threshold = 0.4 at a pixel.  Hence
//   overlap_i  = sum over earlier-kept same-class things j of |B_i & B_j|         (pair counts)
//   A_i        = B_i minus the union of B_k of earlier-kept things                (first kept owner)
// and one counting pass + an O(K^2) greedy + an "owner" pass reproduce the sequential NumPy loop.
//
// Pipeline (all on one stream):
//   select  (1 CTA)   class softmax / keep / order            -> Sel
//   count   (pixels)  |B_i| and pair counts of thing candidates
//   greedy  (1 CTA)   mask_removal decisions                   -> final kept list (stuff.., things..)
//   owner   (pixels)  owner[pixel] = first kept thing with prob >= 0.4   (uint16, 0xFFFF = none)
//   repeat <= max_iters: argmax (pixels) -> ids + areas ; filter (last block of the pass) -> drop area <= small_area
//   (greedy and filter run in the LAST block of the preceding pixel pass: 3 + max_iters launches instead of 16; when the
//    launched iterations do not reach the fixed point -- the reference loops without bound, :761-792 -- the id map is a
//    255 sentinel, meta[3] = 0, and slotvps_panoptic_fuse_resume continues from the device state)
//   relabel (pixels)  ids -> int64 panoptic labels (stuff class | 11 + thing rank), meta
// Bilinear sampling follows F.interpolate(mode="bilinear", align_corners=False) to size (H,W).
#pragma once
#include "common.cuh"

namespace slotvps {

constexpr int FUSE_MAXN = 512;

struct FuseState {                 // lives in the workspace (device)
  int K;                           // kept by score/class filter
  int n_stuff, n_cand;             // stuff count, thing candidates (K = n_stuff + n_cand)
  int Kl;                          // after mask_removal: n_stuff + kept things
  int n_things_kept;
  int iters, converged;
  unsigned int ticket;             // blocks that finished the current pixel pass (the last one runs the serial step)
  int ord[FUSE_MAXN];              // slot index per position: stuff (score desc), then thing candidates (score desc)
  int cls[FUSE_MAXN];              // class per position (same indexing as ord)
  float score[FUSE_MAXN];
  int list[FUSE_MAXN];             // final list after greedy: positions into ord[] (stuff.., kept things..)
  int thing_rank[FUSE_MAXN];       // candidate index -> rank among kept things, or -1
  int active[FUSE_MAXN];           // per final-list entry
  unsigned int area[FUSE_MAXN];    // per final-list entry (current iteration)
  unsigned int cntB[FUSE_MAXN];    // per thing candidate
  int lut[FUSE_MAXN];              // final-list entry -> panoptic label (valid once converged)
};

struct Bilin {
  int i00, i01, i10, i11;
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Bilin bilin_setup(int y, int x, int h, int w, float sy_scale, float sx_scale) {
  float sy = fmaxf(sy_scale * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(sx_scale * (x + 0.5f) - 0.5f, 0.f);
  int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
  int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  float ly = sy - y0, lx = sx - x0;
  Bilin b;
  b.i00 = y0 * w + x0; b.i01 = y0 * w + x1; b.i10 = y1 * w + x0; b.i11 = y1 * w + x1;
  b.w00 = (1.f - ly) * (1.f - lx); b.w01 = (1.f - ly) * lx; b.w10 = ly * (1.f - lx); b.w11 = ly * lx;
  return b;
}
// torch: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11); evaluated with separated weights to keep
// the same association (rounding differs by <= 1-2 ulp, inside the stated logit tolerance)
__device__ __forceinline__ float bilin_eval(const float* __restrict__ m, const Bilin& b, float ly1, float ly0, float lx1, float lx0) {
  return ly0 * (lx0 * __ldg(m + b.i00) + lx1 * __ldg(m + b.i01)) + ly1 * (lx0 * __ldg(m + b.i10) + lx1 * __ldg(m + b.i11));
}
struct Samp {
  int i00, i01, i10, i11;
  float ly0, ly1, lx0, lx1;
  __device__ __forceinline__ float at(const float* __restrict__ m) const {
    return ly0 * (lx0 * __ldg(m + i00) + lx1 * __ldg(m + i01)) + ly1 * (lx0 * __ldg(m + i10) + lx1 * __ldg(m + i11));
  }
};
__device__ __forceinline__ Samp samp_setup(int y, int x, int h, int w, float sy_scale, float sx_scale, bool same) {
  Samp s;
  if (same) { s.i00 = s.i01 = s.i10 = s.i11 = y * w + x; s.ly0 = 1.f; s.ly1 = 0.f; s.lx0 = 1.f; s.lx1 = 0.f; return s; }
  float sy = fmaxf(sy_scale * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(sx_scale * (x + 0.5f) - 0.5f, 0.f);
  int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
  int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  s.ly1 = sy - y0; s.ly0 = 1.f - s.ly1; s.lx1 = sx - x0; s.lx0 = 1.f - s.lx1;
  s.i00 = y0 * w + x0; s.i01 = y0 * w + x1; s.i10 = y1 * w + x0; s.i11 = y1 * w + x1;
  return s;
}

// ---- select: class softmax, keep filter, ordering ---------------------------------------------
__global__ void __launch_bounds__(FUSE_MAXN) fuse_select_kernel(const float* __restrict__ logits, int N, int ncls, int stuff_num,
                                                                float thr, int has_no_object, FuseState* __restrict__ st,
                                                                unsigned int* __restrict__ pair, int* __restrict__ meta) {
  for (int j = threadIdx.x; j < N * N; j += blockDim.x) pair[j] = 0u;         // pair-overlap counters and the result record start clean
  for (int j = threadIdx.x; j < 4 + 3 * N; j += blockDim.x) meta[j] = 0;
  __shared__ float s_score[FUSE_MAXN];
  __shared__ int s_cls[FUSE_MAXN];
  __shared__ int s_keep[FUSE_MAXN];
  __shared__ int s_code[FUSE_MAXN];
  __shared__ int s_cnt[2];
  const int i = threadIdx.x;
  if (i < 2) s_cnt[i] = 0;
  float sc = 0.f;
  int cl = 0, keep = 0;
  if (i < N) {
    const float* l = logits + (long)i * ncls;
    float mx = l[0];
    cl = 0;
    for (int j = 1; j < ncls; ++j) if (l[j] > mx) { mx = l[j]; cl = j; }
    float sum = 0.f;
    for (int j = 0; j < ncls; ++j) sum += expf(l[j] - mx);
    sc = 1.f / sum;                                   // softmax value of the arg-max class (exp(0)/sum)
    // vps_temporal_slots.py:688-693: the no-object test applies only when the logits carry that column (width == num_classes)
    keep = (!has_no_object || cl != ncls - 1) && (sc > thr);
  }
  s_score[i] = sc; s_cls[i] = cl; s_keep[i] = keep;
  __syncthreads();
  if (keep) {
    const bool stuff = cl <= stuff_num - 1;
    int rank = 0;
    for (int j = 0; j < N; ++j) {
      if (!s_keep[j] || j == i) continue;
      if ((s_cls[j] <= stuff_num - 1) != stuff) continue;
      // np.argsort(score)[::-1]: descending score, equal scores -> higher index first
      if (s_score[j] > sc || (s_score[j] == sc && j > i)) ++rank;
    }
    atomicAdd(&s_cnt[stuff ? 0 : 1], 1);
    s_code[i] = rank + (stuff ? 0 : 100000);     // separate array: other threads are still reading s_keep
  }
  __syncthreads();
  const int ns = s_cnt[0], nc = s_cnt[1];
  if (keep) {
    int code = s_code[i];
    int pos = code >= 100000 ? ns + (code - 100000) : code;
    st->ord[pos] = i; st->cls[pos] = cl; st->score[pos] = sc;
  }
  if (i == 0) { st->K = ns + nc; st->n_stuff = ns; st->n_cand = nc; st->Kl = 0; st->n_things_kept = 0; st->iters = 0; st->converged = 0; st->ticket = 0u; }
  for (int j = i; j < FUSE_MAXN; j += blockDim.x) { st->cntB[j] = 0; st->area[j] = 0; st->active[j] = 0; }
}

// ---- greedy: the sequential decisions of mask_removal --------------------------------------------
// One warp: candidate i is decided in order; lanes sum its overlaps with earlier survivors.
__device__ __forceinline__ void fuse_greedy_warp(FuseState* st, const unsigned int* pair, long HW, double frac_thr, int* s_rank, int* s_cls) {
  const int lane = threadIdx.x & 31;
  const int ns = st->n_stuff, nc = st->n_cand;
  for (int j = lane; j < nc; j += 32) { s_rank[j] = -1; s_cls[j] = st->cls[ns + j]; }
  for (int k = lane; k < ns; k += 32) { st->list[k] = k; st->active[k] = 1; }
  __syncwarp();
  int nk = 0;
  for (int i = 0; i < nc; ++i) {
    const unsigned int b = __ldcg(&st->cntB[i]);
    unsigned int ov = 0;
    for (int j = lane; j < i; j += 32)
      if (s_rank[j] >= 0 && s_cls[j] == s_cls[i]) ov += __ldcg(&pair[i * nc + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ov += __shfl_xor_sync(0xffffffffu, ov, o);
    bool keep = !(b == 0 || (long)b == HW);                      // constant binarisation (:620)
    if (keep && (double)ov / (double)(float)b > frac_thr) keep = false;   // int64/float32 -> float64 (:621-622)
    if (keep) {
      if (lane == 0) { s_rank[i] = nk; st->list[ns + nk] = ns + i; st->active[ns + nk] = 1; }
      ++nk;
    }
    __syncwarp();
  }
  for (int j = lane; j < nc; j += 32) st->thing_rank[j] = s_rank[j];
  if (lane == 0) { st->Kl = ns + nk; st->n_things_kept = nk; }
}

// "last block" hand-over: every block of a pixel pass publishes its global atomics, takes a ticket, and the block that
// draws the last one runs the serial step that used to be a separate single-CTA launch (count -> greedy, argmax -> filter)
__device__ __forceinline__ bool fuse_last_block(FuseState* st) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&st->ticket, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) st->ticket = 0u;                                  // ready for the next pass (stream order separates the passes)
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// ---- count: |B_i| and pair overlaps of thing candidates ------------------------------------------
// pair [n_cand][n_cand] (global, zeroed by the caller); one thread per output pixel.
__global__ void __launch_bounds__(256) fuse_count_kernel(const float* __restrict__ masks, int h, int w, int H, int W,
                                                         float pix_thr, double frac_thr, FuseState* __restrict__ st, unsigned int* __restrict__ pair, unsigned int* __restrict__ cand) {
  __shared__ unsigned int s_cnt[FUSE_MAXN];
  __shared__ int s_gcls[FUSE_MAXN];
  const int K = st->K, ns = st->n_stuff, nc = st->n_cand;
  for (int j = threadIdx.x; j < nc; j += 256) s_cnt[j] = 0;
  __syncthreads();
  const long P = (long)h * w;
  const float sys = (float)h / H, sxs = (float)w / W;
  const bool same = (h == H && w == W);
  for (long pix = (long)blockIdx.x * 256 + threadIdx.x; pix < (long)H * W; pix += (long)gridDim.x * 256) {
    const int y = (int)(pix / W), x = (int)(pix % W);
    const Samp sp = samp_setup(y, x, h, w, sys, sxs, same);
    // running-maximum softmax and log-domain threshold, the same arithmetic as fuse_count4_kernel
    float mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < K; ++k) {
      const float v = sp.at(masks + (long)st->ord[k] * P);
      const float dlt = v - mx;
      const float ex = __expf(-fabsf(dlt));
      sum = dlt > 0.f ? fmaf(sum, ex, 1.f) : sum + ex;
      mx = fmaxf(mx, v);
    }
    const float cut = mx + logf(pix_thr * sum);
    int c1 = -1, c2 = -1;
    for (int k = ns; k < K; ++k) {
      if (sp.at(masks + (long)st->ord[k] * P) >= cut) { if (c1 < 0) c1 = k - ns; else if (c2 < 0) c2 = k - ns; }
    }
    if (c1 >= 0) atomicAdd(&s_cnt[c1], 1u);
    if (c2 >= 0) { atomicAdd(&s_cnt[c2], 1u); atomicAdd(&pair[c1 * nc + c2], 1u); atomicAdd(&pair[c2 * nc + c1], 1u); }
    cand[pix] = (unsigned int)(c1 & 0xFFFF) | ((unsigned int)(c2 & 0xFFFF) << 16);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nc; j += 256) if (s_cnt[j]) atomicAdd(&st->cntB[j], s_cnt[j]);
  if (fuse_last_block(st) && threadIdx.x < 32) fuse_greedy_warp(st, pair, (long)H * W, frac_thr, reinterpret_cast<int*>(s_cnt), s_gcls);
}

// owner of a pixel = rank (among surviving things) of the first surviving candidate, or 0xFFFF
__device__ __forceinline__ int owner_of(unsigned int cd, const int* __restrict__ rank) {
  const int c1 = cd & 0xFFFF, c2 = cd >> 16;
  if (c1 != 0xFFFF && rank[c1] >= 0) return rank[c1];
  if (c2 != 0xFFFF && rank[c2] >= 0) return rank[c2];
  return 0xFFFF;
}

// ---- x4 specialisation (H == 4h, W == 4w): one thread per 4x4 output block shares 3x3 source taps per slot.
// The per-pixel arithmetic (tap indices, weights, association) is IDENTICAL to Samp::at, so results are
// bit-identical to the generic kernels.
struct Blk4 {
  int r[3], c[3];            // clamped source rows / cols  (i-1, i, i+1), (j-1, j, j+1)
  float ly0[4], ly1[4], lx0[4], lx1[4];
  // Output row 4i+a taps source rows (i-1, i) for a < 2 and (i, i+1) for a >= 2 (same for columns).  At the
  // borders the generic formula clamps the coordinate / the second tap; with the clamped rows r[] the static
  // tap pairs then carry the generic weights onto the same texels, so every value is bit-identical.
  __device__ __forceinline__ void setup(int i, int j, int h, int w) {
#pragma unroll
    for (int m = 0; m < 3; ++m) { r[m] = min(max(i - 1 + m, 0), h - 1); c[m] = min(max(j - 1 + m, 0), w - 1); }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float sy = fmaxf(0.25f * (4 * i + a + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.25f * (4 * j + a + 0.5f) - 0.5f, 0.f);
      int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
      ly1[a] = sy - y0; ly0[a] = 1.f - ly1[a]; lx1[a] = sx - x0; lx0[a] = 1.f - lx1[a];
    }
  }
  // the 3x3 source taps of the block; returns their maximum (every output value of the block is a convex combination of them)
  __device__ __forceinline__ float taps(const float* __restrict__ m, int w, float (*t)[3]) const {
    float mx = -INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) { t[a][b] = __ldg(m + r[a] * w + c[b]); mx = fmaxf(mx, t[a][b]); }
    return mx;
  }
  // v[a*4+b] = value at output pixel (4i+a, 4j+b)
  __device__ __forceinline__ void eval(const float* __restrict__ m, int w, float* v) const {
    float t[3][3];
    taps(m, w, t);
    interp(t, v);
  }
  __device__ __forceinline__ void interp(const float (*t)[3], float* v) const {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ra = a < 2 ? 0 : 1, cb = b < 2 ? 0 : 1;
        v[a * 4 + b] = ly0[a] * (lx0[b] * t[ra][cb] + lx1[b] * t[ra][cb + 1]) + ly1[a] * (lx0[b] * t[ra + 1][cb] + lx1[b] * t[ra + 1][cb + 1]);
      }
  }
  __device__ __forceinline__ float eval1(const float* __restrict__ m, int w, int a, int b) const {
    const int ra = a < 2 ? 0 : 1, cb = b < 2 ? 0 : 1;
    return ly0[a] * (lx0[b] * __ldg(m + r[ra] * w + c[cb]) + lx1[b] * __ldg(m + r[ra] * w + c[cb + 1])) +
           ly1[a] * (lx0[b] * __ldg(m + r[ra + 1] * w + c[cb]) + lx1[b] * __ldg(m + r[ra + 1] * w + c[cb + 1]));
  }
};

__global__ void __launch_bounds__(256) fuse_count4_kernel(const float* __restrict__ masks, int h, int w, float pix_thr, double frac_thr,
                                                          FuseState* __restrict__ st, unsigned int* __restrict__ pair,
                                                          unsigned int* __restrict__ cand) {
  __shared__ unsigned int s_cnt[FUSE_MAXN];
  __shared__ int s_gcls[FUSE_MAXN];
  __shared__ int s_ord[FUSE_MAXN];
  const int K = st->K, ns = st->n_stuff, nc = st->n_cand;
  for (int j = threadIdx.x; j < FUSE_MAXN; j += 256) { s_cnt[j] = 0; s_ord[j] = j < K ? st->ord[j] : 0; }
  __syncthreads();
  const long P = (long)h * w;
  const int W = 4 * w;
  const float screen_cut = pix_thr > 0.f ? __logf(pix_thr) - 0.05f : -INFINITY;     // margin >> fp32 rounding of the bound
  for (long blk = (long)blockIdx.x * 256 + threadIdx.x; blk < P; blk += (long)gridDim.x * 256) {
    const int i = (int)(blk / w), j = (int)(blk % w);
    Blk4 b4;
    b4.setup(i, j, h, w);
    // Conservative screen: every output value of the block is a convex combination of the slot's 3x3 source taps,
    // so p_k <= exp(max tap_k) / sum_j exp(min tap_j).  If that bound stays below pixel_threshold for every thing,
    // no pixel of the block has a candidate and the three exact passes are skipped (stuff-only regions).
    float mlow;
    {
      float umax = -INFINITY, m = -INFINITY, ssum = 0.f;
      for (int k = 0; k < K; ++k) {
        const float* mk = masks + (long)s_ord[k] * P;
        float tmax = -INFINITY, tmin = INFINITY;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) { const float tv = __ldg(mk + b4.r[a] * w + b4.c[b]); tmax = fmaxf(tmax, tv); tmin = fminf(tmin, tv); }
        if (k >= ns) umax = fmaxf(umax, tmax);
        if (tmin > m) { ssum = ssum * __expf(m - tmin) + 1.f; m = tmin; } else ssum += __expf(tmin - m);
      }
      mlow = m;
      if (ns == K || umax - (m + __logf(ssum)) < screen_cut) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
          *reinterpret_cast<uint4*>(cand + (long)(4 * i + a) * W + 4 * j) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        continue;
      }
    }
    // softmax over the kept slots in one sweep (running maximum, sum rescaled when it moves: one exp per value), then the
    // threshold test in the log domain: p_k >= thr  <=>  v_k >= max + log(thr * sum).  Against the reference's
    // exp / divide / compare this moves the decision only for pixels whose p is within ~1e-6 (relative) of the threshold.
    float mx[16], sum[16], v[16], t[3][3];
#pragma unroll
    for (int e = 0; e < 16; ++e) { mx[e] = -INFINITY; sum[e] = 0.f; }
    // A slot whose largest tap lies 20 below `mlow` (= the largest of the slots' smallest taps, a lower bound of the maximum at
    // every pixel of the block) adds less than e^-20 = 2e-9 of the sum at every pixel -- 27 such slots stay below half an ulp of
    // the fp32 sum -- so its interpolation and exponentials are skipped; inside an object that is most of the kept slots.
    const float skip_below = mlow - 20.f;
    for (int k = 0; k < K; ++k) {
      if (b4.taps(masks + (long)s_ord[k] * P, w, t) < skip_below) continue;
      b4.interp(t, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float dlt = v[e] - mx[e];
        const float ex = __expf(-fabsf(dlt));
        sum[e] = dlt > 0.f ? fmaf(sum[e], ex, 1.f) : sum[e] + ex;
        mx[e] = fmaxf(mx[e], v[e]);
      }
    }
    float cutmin = INFINITY;
#pragma unroll
    for (int e = 0; e < 16; ++e) { mx[e] += logf(pix_thr * sum[e]); cutmin = fminf(cutmin, mx[e]); }     // the cut: -inf when pix_thr == 0
    unsigned int cd[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) cd[e] = 0xFFFFFFFFu;
    for (int k = ns; k < K; ++k) {
      if (b4.taps(masks + (long)s_ord[k] * P, w, t) < cutmin) continue;       // exact: no value of the block can reach its cut
      b4.interp(t, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        if (v[e] >= mx[e]) {
          const unsigned int c = (unsigned int)(k - ns);
          if ((cd[e] & 0xFFFF) == 0xFFFF) cd[e] = (cd[e] & 0xFFFF0000u) | c;
          else if ((cd[e] >> 16) == 0xFFFF) cd[e] = (cd[e] & 0xFFFFu) | (c << 16);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      *reinterpret_cast<uint4*>(cand + (long)(4 * i + a) * W + 4 * j) = make_uint4(cd[a * 4], cd[a * 4 + 1], cd[a * 4 + 2], cd[a * 4 + 3]);
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const unsigned int c = cd[a * 4 + bb];
        const int c1 = c & 0xFFFF, c2 = c >> 16;
        if (c1 != 0xFFFF) atomicAdd(&s_cnt[c1], 1u);
        if (c2 != 0xFFFF) { atomicAdd(&s_cnt[c2], 1u); atomicAdd(&pair[c1 * nc + c2], 1u); atomicAdd(&pair[c2 * nc + c1], 1u); }
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nc; j += 256) if (s_cnt[j]) atomicAdd(&st->cntB[j], s_cnt[j]);
  if (fuse_last_block(st) && threadIdx.x < 32) fuse_greedy_warp(st, pair, (long)16 * h * w, frac_thr, reinterpret_cast<int*>(s_cnt), s_gcls);
}

// ---- filter: merge duplicate stuff classes (first call only), drop area <= small_area; on
// convergence build the label LUT of the inline fusion (vps_temporal_slots.py:420-435) and meta.
// Runs in the LAST block of an argmax pass (all 256 threads), entries strided over the threads. ----
struct FilterArgs { int stuff_num; unsigned int small_area; int N; int* meta; };
__device__ __forceinline__ void fuse_filter_block(FuseState* st, const FilterArgs fa) {
  __shared__ unsigned int f_area[FUSE_MAXN];
  __shared__ int f_act[FUSE_MAXN], f_cls[FUSE_MAXN], f_first[FUSE_MAXN], f_comp[FUSE_MAXN], f_lut[FUSE_MAXN];
  __shared__ int f_removed;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int Kl = st->Kl, ns = st->n_stuff, it = st->iters;
  if (tid == 0) f_removed = 0;
  for (int e = tid; e < Kl; e += nt) { f_area[e] = __ldcg(&st->area[e]); f_act[e] = st->active[e]; f_cls[e] = st->cls[st->list[e]]; }
  __syncthreads();
  if (it == 0) {                                                  // dedup=True only on the first call (:758)
    for (int e = tid; e < Kl; e += nt) {
      int first = e;
      if (e < ns && f_act[e])
        for (int f = 0; f < e; ++f) if (f_act[f] && f_cls[f] == f_cls[e]) { first = f; break; }
      f_first[e] = first;
    }
    __syncthreads();
    for (int e = tid; e < ns; e += nt) {
      if (f_act[e] && f_first[e] == e) {
        unsigned int a = f_area[e];
        for (int f = e + 1; f < ns; ++f) if (f_act[f] && f_first[f] == e) a += f_area[f];
        f_comp[e] = (int)a;                                       // staged: other threads still read f_area of merged entries
      }
    }
    __syncthreads();
    for (int e = tid; e < ns; e += nt) {
      if (!f_act[e]) continue;
      f_area[e] = f_first[e] == e ? (unsigned int)f_comp[e] : 0u;
    }
    __syncthreads();
  }
  for (int e = tid; e < Kl; e += nt)
    if (f_act[e] && f_area[e] <= fa.small_area) { f_act[e] = 0; st->active[e] = 0; atomicAdd(&f_removed, 1); }
  __syncthreads();
  if (f_removed != 0) {
    for (int e = tid; e < Kl; e += nt) st->area[e] = 0;
    if (tid == 0) st->iters = it + 1;
    return;
  }
  // converged: every active entry is present (area > small_area) in the id map of this iteration
  if (tid == 0) {
    int n_all = 0, n_inst = 0;
    for (int f = 0; f < Kl; ++f) { f_comp[f] = f_act[f] ? n_all : -1; if (f_act[f]) { f_first[n_all] = f; ++n_all; if (f >= ns) ++n_inst; } }
    // present ids ascending = compacted ids 0..n_all-1 restricted to area > 0; scan from the top
    int npresent = 0;
    for (int f = 0; f < Kl; ++f) if (f_act[f] && f_area[f] > 0) ++npresent;
    int count = n_inst, i = npresent - 1;
    for (int f = Kl - 1; f >= 0; --f) {
      f_lut[f] = 0;
      if (!(f_act[f] && f_area[f] > 0)) continue;
      if (f_comp[f] >= n_all - n_inst) { f_lut[f] = fa.stuff_num + count - 1; --count; }
      else f_lut[f] = f_cls[f_first[i]];      // semantic_labels[i], i = position in the unique-id list (:433)
      --i;
    }
    int* meta = fa.meta;
    meta[0] = n_all; meta[1] = n_inst; meta[2] = it + 1; meta[3] = 1;
    for (int c = 0; c < n_all; ++c) {
      const int f = f_first[c];
      meta[4 + c] = st->ord[st->list[f]];
      meta[4 + fa.N + c] = f_cls[f];
      meta[4 + 2 * fa.N + c] = __float_as_int(st->score[st->list[f]]);
    }
    st->iters = it + 1; st->converged = 1;
  }
  __syncthreads();
  for (int e = tid; e < Kl; e += nt) st->lut[e] = f_lut[e];
}

__global__ void __launch_bounds__(256) fuse_argmax4_kernel(const float* __restrict__ masks, int h, int w, FuseState* __restrict__ st,
                                                           const unsigned int* __restrict__ cand, unsigned short* __restrict__ ids, const FilterArgs fa) {
  if (st->converged) return;
  __shared__ unsigned int s_area[FUSE_MAXN];
  __shared__ int s_rank[FUSE_MAXN], s_slot[FUSE_MAXN], s_act[FUSE_MAXN];
  __shared__ int s_first_thing, s_second_thing;
  const int Kl = st->Kl, ns = st->n_stuff, nc = st->n_cand;
  for (int j = threadIdx.x; j < FUSE_MAXN; j += 256) {
    s_area[j] = 0;
    s_rank[j] = j < nc ? st->thing_rank[j] : -1;
    s_slot[j] = j < Kl ? st->ord[st->list[j]] : 0;
    s_act[j] = j < Kl ? st->active[j] : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int f = -1, s2 = -1;
    for (int e = ns; e < Kl; ++e) if (s_act[e]) { if (f < 0) f = e; else if (s2 < 0) { s2 = e; break; } }
    s_first_thing = f; s_second_thing = s2;
  }
  __syncthreads();
  const int first_thing = s_first_thing, second_thing = s_second_thing;
  const long P = (long)h * w;
  const int W = 4 * w;
  for (long blk = (long)blockIdx.x * 256 + threadIdx.x; blk < P; blk += (long)gridDim.x * 256) {
    const int i = (int)(blk / w), j = (int)(blk % w);
    Blk4 b4;
    b4.setup(i, j, h, w);
    float best[16], v[16];
    int bi[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) { best[e] = -INFINITY; bi[e] = -1; }
    for (int s = 0; s < ns; ++s) {
      if (!s_act[s]) continue;
      b4.eval(masks + (long)s_slot[s] * P, w, v);
#pragma unroll
      for (int e = 0; e < 16; ++e) if (v[e] > best[e]) { best[e] = v[e]; bi[e] = s; }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const uint4 cq = *reinterpret_cast<const uint4*>(cand + (long)(4 * i + a) * W + 4 * j);
      const unsigned int cds[4] = {cq.x, cq.y, cq.z, cq.w};
      unsigned short out4[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int e = a * 4 + bb;
        float bst = best[e];
        int b = bi[e];
        if (first_thing >= 0) {
          const int own = owner_of(cds[bb], s_rank);
          const int oe = own == 0xFFFF ? -1 : ns + own;
          const bool own_active = oe >= 0 && s_act[oe];
          int ze = !own_active ? first_thing : ((first_thing != oe) ? first_thing : second_thing);
          float tv = -INFINITY;
          int te = -1;
          if (own_active) {
            tv = b4.eval1(masks + (long)s_slot[oe] * P, w, a, bb);
            te = oe;
          }
          if (ze >= 0 && (0.f > tv || (0.f == tv && ze < te))) { tv = 0.f; te = ze; }
          if (tv > bst) { bst = tv; b = te; }
        }
        out4[bb] = (unsigned short)b;
        if (b >= 0) atomicAdd(&s_area[b], 1u);
      }
      *reinterpret_cast<uint2*>(ids + (long)(4 * i + a) * W + 4 * j) =
          make_uint2((unsigned int)out4[0] | ((unsigned int)out4[1] << 16), (unsigned int)out4[2] | ((unsigned int)out4[3] << 16));
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Kl; j += 256) if (s_area[j]) atomicAdd(&st->area[j], s_area[j]);
  if (fuse_last_block(st)) fuse_filter_block(st, fa);
}

// ---- argmax over the active kept slots of the masked logits ---------------------------------------
// ids[pixel] = entry of the final list (uncompacted); lowest entry wins ties (torch argmax).
__global__ void __launch_bounds__(256) fuse_argmax_kernel(const float* __restrict__ masks, int h, int w, int H, int W,
                                                          FuseState* __restrict__ st, const unsigned int* __restrict__ cand,
                                                          unsigned short* __restrict__ ids, const FilterArgs fa) {
  if (st->converged) return;
  __shared__ unsigned int s_area[FUSE_MAXN];
  __shared__ int s_rank[FUSE_MAXN];
  __shared__ int s_first_thing, s_second_thing;
  for (int j = threadIdx.x; j < FUSE_MAXN; j += 256) s_rank[j] = j < st->n_cand ? st->thing_rank[j] : -1;
  const int Kl = st->Kl, ns = st->n_stuff;
  for (int j = threadIdx.x; j < Kl; j += 256) s_area[j] = 0;
  if (threadIdx.x == 0) {
    int f = -1, s2 = -1;
    for (int e = ns; e < Kl; ++e) if (st->active[e]) { if (f < 0) f = e; else if (s2 < 0) { s2 = e; break; } }
    s_first_thing = f; s_second_thing = s2;
  }
  __syncthreads();
  const int first_thing = s_first_thing, second_thing = s_second_thing;
  const long P = (long)h * w;
  const float sys = (float)h / H, sxs = (float)w / W;
  const bool same = (h == H && w == W);
  for (long pix = (long)blockIdx.x * 256 + threadIdx.x; pix < (long)H * W; pix += (long)gridDim.x * 256) {
    const int y = (int)(pix / W), x = (int)(pix % W);
    const Samp sp = samp_setup(y, x, h, w, sys, sxs, same);
    float best = -INFINITY;
    int bi = -1;
    for (int e = 0; e < ns; ++e) {
      if (!st->active[e]) continue;
      float v = sp.at(masks + (long)st->ord[st->list[e]] * P);
      if (v > best) { best = v; bi = e; }
    }
    if (first_thing >= 0) {
      const int own = owner_of(cand[pix], s_rank);
      const int oe = own == 0xFFFF ? -1 : ns + own;            // entry of the owner (things keep their order)
      const bool own_active = oe >= 0 && st->active[oe];
      // all active things other than the owner carry exactly 0 here (:632-634)
      int ze = -1;                                              // lowest active thing entry holding a zero
      if (!own_active) ze = first_thing;
      else ze = (first_thing != oe) ? first_thing : second_thing;
      float tv = -INFINITY;
      int te = -1;
      if (own_active) { tv = sp.at(masks + (long)st->ord[st->list[oe]] * P); te = oe; }
      if (ze >= 0 && (0.f > tv || (0.f == tv && ze < te))) { tv = 0.f; te = ze; }
      if (tv > best) { best = tv; bi = te; }
    }
    ids[pix] = (unsigned short)bi;
    if (bi >= 0) atomicAdd(&s_area[bi], 1u);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Kl; j += 256) if (s_area[j]) atomicAdd(&st->area[j], s_area[j]);
  if (fuse_last_block(st)) fuse_filter_block(st, fa);
}

// ---- relabel: ids -> int64 panoptic labels -------------------------------------------------------
__global__ void __launch_bounds__(256) fuse_relabel_kernel(const FuseState* __restrict__ st, const unsigned short* __restrict__ ids,
                                                           long HW, long long* __restrict__ out) {
  __shared__ int s_lut[FUSE_MAXN];
  const int Kl = st->Kl;
  for (int j = threadIdx.x; j < Kl; j += 256) s_lut[j] = st->lut[j];
  __syncthreads();
  const bool conv = st->converged != 0;                           // not converged within the launched iterations: sentinel map,
  for (long pix = (long)blockIdx.x * 256 + threadIdx.x; pix < HW; pix += (long)gridDim.x * 256) {      // meta[3] stays 0, the host resumes
    const unsigned short id = ids[pix];
    out[pix] = !conv ? 255 : (id == 0xFFFF ? 0 : s_lut[id]);
  }
}

// ---- optional: materialise Instances.masks (masked logits of the final kept slots) -------------------
__global__ void __launch_bounds__(256) fuse_masks_kernel(const float* __restrict__ masks, int h, int w, int H, int W,
                                                         const FuseState* __restrict__ st, const unsigned int* __restrict__ cand,
                                                         float* __restrict__ out, int cap) {
  __shared__ int s_rank[FUSE_MAXN];
  for (int j = threadIdx.x; j < FUSE_MAXN; j += 256) s_rank[j] = j < st->n_cand ? st->thing_rank[j] : -1;
  __syncthreads();
  const int Kl = st->Kl, ns = st->n_stuff;
  const long P = (long)h * w, HW = (long)H * W;
  const float sys = (float)h / H, sxs = (float)w / W;
  const bool same = (h == H && w == W);
  for (long pix = (long)blockIdx.x * 256 + threadIdx.x; pix < HW; pix += (long)gridDim.x * 256) {
    const int y = (int)(pix / W), x = (int)(pix % W);
    const Samp sp = samp_setup(y, x, h, w, sys, sxs, same);
    const int own = owner_of(cand[pix], s_rank);
    int c = 0;
    for (int e = 0; e < Kl && c < cap; ++e) {
      if (!st->active[e]) continue;
      float v;
      if (e < ns) v = sp.at(masks + (long)st->ord[st->list[e]] * P);
      else v = (own != 0xFFFF && ns + own == e) ? sp.at(masks + (long)st->ord[st->list[e]] * P) : 0.f;
      out[(long)c * HW + pix] = v;
      ++c;
    }
  }
}

}  // namespace slotvps
