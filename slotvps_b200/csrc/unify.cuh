// Consumers of the id map, on device:
//   semantic_argmax_kernel  -- simple_test's semantic prediction (vps_temporal_slots.py:440-451): optional bilinear
//                              resize (align_corners=False), softmax over classes, first index of the max;
//   unify_*                 -- CityscapesVps.get_unified_pan_result (tools/dataset/cityscapes_vps.py:214-302): per
//                              frame, reconcile semantic argmax, panoptic ids, thing classes and object ids into the
//                              H x W x 3 uint8 wire format (semantic label, instance index from 1, object id + 1).
// The reference's NumPy version runs np.unique / boolean-mask passes per instance (O(K * H * W) per frame); here one
// histogram pass over the pixels, one single-block decision kernel that replays the sequential rules on the
// 256 x 32 (id, class) table, and one pass that writes the three channels through 256-entry look-up tables.
#pragma once
#include "common.cuh"
#include "fusion.cuh"

namespace slotvps {
namespace unify {
constexpr int MAXID = 256;     // panoptic ids (stuff label or stuff_num + instance) must be < 256 (they are stored as uint8)
constexpr int MAXSEM = 32;     // semantic classes
struct State {                 // first two words = the `status` record of slotvps_unify_pan_result
  int error;                   // sticky: bit0 id/class out of range, bit1 cls_inds index out of range, bit2 obj_ids too short
  int max_oid;                 // the counter for redundant object ids (:219), persists over the frames of a call
  int pad[2];
};

__global__ void unify_reset_kernel(State* st) { st->max_oid = 100; st->error = 0; }

__global__ void __launch_bounds__(256) unify_hist_kernel(const long long* __restrict__ seg, const long long* __restrict__ pan, long HW,
                                                         unsigned int* __restrict__ hist, State* __restrict__ st) {
  __shared__ unsigned int s_h[MAXID * MAXSEM];
  for (int i = threadIdx.x; i < MAXID * MAXSEM; i += 256) s_h[i] = 0;
  __syncthreads();
  bool bad = false;
  // 8 consecutive pixels per thread: runs of one (id, class) pair cost one shared-memory atomic
  for (long base = ((long)blockIdx.x * 256 + threadIdx.x) * 8; base < HW; base += (long)gridDim.x * 256 * 8) {
    int key = -1;
    unsigned int run = 0;
    for (int e = 0; e < 8 && base + e < HW; ++e) {
      const long long p = pan[base + e], s = seg[base + e];
      if (p < 0 || p >= MAXID || s < 0 || s >= MAXSEM) { bad = true; continue; }
      const int k = (int)p * MAXSEM + (int)s;
      if (k == key) { ++run; continue; }
      if (run) atomicAdd(&s_h[key], run);
      key = k; run = 1;
    }
    if (run) atomicAdd(&s_h[key], run);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MAXID * MAXSEM; i += 256) if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
  if (bad) atomicOr(&st->error, 1);
}

// One block, thread t owns id t for the reductions; thread 0 replays the sequential rules.
// luts: [3][256] uint8 (semantic, instance, object) indexed by the ORIGINAL pan value.
__global__ void __launch_bounds__(MAXID) unify_decide_kernel(const unsigned int* __restrict__ hist, const int* __restrict__ cls_inds,
                                                            int n_inst, const int* __restrict__ obj_ids, int n_obj, int last_stuff,
                                                            unsigned int area_limit, State* __restrict__ st, unsigned char* __restrict__ luts) {
  __shared__ unsigned int s_total[MAXID], s_maxc[MAXID];
  __shared__ int s_major[MAXID];
  __shared__ int s_seg[MAXID], s_ins[MAXID], s_obj[MAXID];
  __shared__ int s_oid[MAXID], s_rev[MAXID];
  const int t = threadIdx.x;
  {
    unsigned int tot = 0, mc = 0;
    int mj = 0;
    for (int s = 0; s < MAXSEM; ++s) {
      const unsigned int c = hist[t * MAXSEM + s];
      tot += c;
      if (c > mc) { mc = c; mj = s; }                 // first maximum = smallest class among ties (np.argmax of np.unique counts)
    }
    s_total[t] = tot; s_maxc[t] = mc; s_major[t] = mj;
    s_seg[t] = t; s_ins[t] = t <= last_stuff ? 0 : t; s_obj[t] = t;      // pan_seg / pan_ins / pan_obj start as copies of pan (:248-257)
    if (obj_ids && t < n_obj) s_oid[t] = obj_ids[t];
  }
  __syncthreads();
  if (t == 0) {
    int err = 0;
    // ---- redundant object ids (:233-244): the LAST holder keeps the id, earlier ones get fresh ids from max_oid ----
    if (obj_ids && n_obj > 0) {
      int max_oid = st->max_oid;
      for (int i = 0; i < n_obj; ++i) s_rev[i] = s_oid[n_obj - 1 - i];
      long long last = -(1LL << 40);
      while (true) {                                   // redundant ids in ascending order (np.unique)
        long long red = (1LL << 40);
        for (int i = 0; i < n_obj; ++i) if (s_oid[i] > last && s_oid[i] < red) red = s_oid[i];
        if (red == (1LL << 40)) break;
        last = red;
        int n = 0;
        for (int i = 0; i < n_obj; ++i) n += s_oid[i] == (int)red;
        if (n < 2) continue;
        int j = 0;                                     // replacement values: red, max_oid, max_oid + 1, ...
        for (int i = 0; i < n_obj && j < n; ++i)
          if (s_rev[i] == (int)red) { if (j > 0) s_rev[i] = max_oid++; ++j; }
      }
      for (int i = 0; i < n_obj; ++i) s_oid[i] = s_rev[n_obj - 1 - i];
      st->max_oid = max_oid;
    }
    if (n_inst == 0) {
      // every id above the stuff range collapses to 255 (:251-252, :262-265): semantic 255, no instance, object channel 255
      for (int i = last_stuff + 1; i < MAXID; ++i) { s_seg[i] = 255; s_ins[i] = 0; s_obj[i] = 255; }
    } else {
      int idx = 0;
      for (int i = last_stuff + 1; i < MAXID; ++i) {
        if (s_total[i] == 0) continue;                 // ids_ins = the ids present in the map, ascending (:256-257)
        const int ci = i - last_stuff - 1;
        if (ci >= n_inst) { err |= 2; ++idx; continue; }
        const int inst_cls = cls_inds[ci] + last_stuff;
        const int major = s_major[i];
        // np.max(cnt) / np.sum(cnt) >= 0.5 in float64; exact for counts below 2^52
        const bool outvoted = major != inst_cls && 2ULL * s_maxc[i] >= (unsigned long long)s_total[i] && major <= last_stuff;
        if (outvoted) { s_seg[i] = major; s_ins[i] = 0; s_obj[i] = 0; }
        else {
          s_seg[i] = inst_cls; s_ins[i] = idx + 1;
          if (obj_ids) { if (idx < n_obj) s_obj[i] = s_oid[idx] + 1; else err |= 4; }
        }
        ++idx;
      }
    }
    // ---- stuff classes below the area limit become 255 (:289-294); areas are those of the FINAL semantic map ----
    for (int v = 0; v <= last_stuff && v < MAXID; ++v) {
      unsigned long long area = 0;
      for (int i = 0; i < MAXID; ++i) if (s_seg[i] == v) area += s_total[i];
      if (area > 0 && area < area_limit)
        for (int i = 0; i < MAXID; ++i) if (s_seg[i] == v && s_total[i] > 0) s_seg[i] = 255;
    }
    if (err) atomicOr(&st->error, err);
  }
  __syncthreads();
  luts[t] = (unsigned char)s_seg[t];                   // astype(uint8): wraps modulo 256
  luts[MAXID + t] = (unsigned char)s_ins[t];
  luts[2 * MAXID + t] = (unsigned char)s_obj[t];
}

__global__ void __launch_bounds__(256) unify_write_kernel(const long long* __restrict__ pan, long HW, const unsigned char* __restrict__ luts,
                                                          unsigned char* __restrict__ out) {
  __shared__ unsigned char s_l[3 * MAXID];
  for (int i = threadIdx.x; i < 3 * MAXID; i += 256) s_l[i] = luts[i];
  __syncthreads();
  for (long q = (long)blockIdx.x * 256 + threadIdx.x; q * 4 < HW; q += (long)gridDim.x * 256) {
    const long p0 = q * 4;
    if (p0 + 4 <= HW) {                                // 4 pixels -> 12 bytes = three aligned 32-bit words
      unsigned char b[12];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int id = (int)(pan[p0 + e] & 0xFF);
        b[3 * e] = s_l[id]; b[3 * e + 1] = s_l[MAXID + id]; b[3 * e + 2] = s_l[2 * MAXID + id];
      }
      unsigned int* o = reinterpret_cast<unsigned int*>(out + p0 * 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) o[k] = b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | ((unsigned int)b[4 * k + 3] << 24);
    } else {
      for (long p = p0; p < HW; ++p) {
        const int id = (int)(pan[p] & 0xFF);
        out[p * 3] = s_l[id]; out[p * 3 + 1] = s_l[MAXID + id]; out[p * 3 + 2] = s_l[2 * MAXID + id];
      }
    }
  }
}

// fcn_output [Cs,h,w] -> out [H,W] int64
__global__ void __launch_bounds__(256) semantic_argmax_kernel(const float* __restrict__ x, int Cs, int h, int w, int H, int W,
                                                              long long* __restrict__ out) {
  const long P = (long)h * w;
  const float sys = (float)h / H, sxs = (float)w / W;
  const bool same = (h == H && w == W);
  for (long pix = (long)blockIdx.x * 256 + threadIdx.x; pix < (long)H * W; pix += (long)gridDim.x * 256) {
    const int y = (int)(pix / W), xx = (int)(pix % W);
    const Samp sp = samp_setup(y, xx, h, w, sys, sxs, same);
    float v[MAXSEM];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXSEM; ++c) if (c < Cs) { v[c] = sp.at(x + (long)c * P); mx = fmaxf(mx, v[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < MAXSEM; ++c) if (c < Cs) { v[c] = expf(v[c] - mx); sum += v[c]; }
    float best = -1.f;
    int bi = 0;
#pragma unroll
    for (int c = 0; c < MAXSEM; ++c) if (c < Cs) { const float p = v[c] / sum; if (p > best) { best = p; bi = c; } }
    out[pix] = bi;
  }
}

}  // namespace unify
}  // namespace slotvps
