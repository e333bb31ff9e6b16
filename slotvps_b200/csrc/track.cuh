// On-device tracker: SimpleTrackHead (simple_track_head.py:58-92: `num_fcs` Linear layers
// with ReLU between, correlation with the bank, all-zero "new object" column) and the greedy id assignment of
// simple_test (vps_temporal_slots.py:322-409).  The object bank (prev_instances.output_embedding: RAW slot
// embeddings, row = object id) and its length live in device memory, so a video is processed without a host
// round trip per frame.  Everything here is tiny (K <= N slots, bank <= capacity rows): three launches per frame.
#pragma once
#include "common.cuh"

namespace slotvps {
namespace track {

struct State {              // header of the device-side tracker state; bank rows follow at +256 bytes
  int count;                // bank rows in use (= len(prev_instances))
  int started;              // 0 until the first frame of the video has been absorbed (:335-342)
  int overflow;             // set when an append was dropped because the bank was full
  int pad;
};
constexpr int ROWS = 4;     // rows of the FC stack per block

// y = fc_{L-1}(relu(... relu(fc_0(x)))) for the rows of one operand.
//   src rows are taken through `index` when given (kept slots of the current frame), else row r of `src`.
//   n_ptr: device count of valid rows (meta[0] or State::count).  W [L][C][C] as torch stores nn.Linear ([out][in]).
__global__ void __launch_bounds__(256) track_fc_kernel(const float* __restrict__ src, const int* __restrict__ index,
                                                       const int* __restrict__ n_ptr, int n_host,
                                                       const float* __restrict__ W, const float* __restrict__ B, int L,
                                                       float* __restrict__ dst) {
  __shared__ float s_x[2][ROWS][C];
  const int n = n_ptr ? min(*n_ptr, n_host) : n_host;
  const int r0 = blockIdx.x * ROWS;
  if (r0 >= n) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int r = 0; r < ROWS; ++r) {
    const int row = r0 + r;
    s_x[0][r][tid] = row < n ? src[(long)(index ? index[row] : row) * C + tid] : 0.f;
  }
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < L; ++l) {
    const float* Wl = W + (long)l * C * C;
    for (int j = warp; j < C; j += 8) {             // one warp per output channel, lanes over the input channels
      float acc[ROWS] = {0.f, 0.f, 0.f, 0.f};
      const float* wr = Wl + (long)j * C;
#pragma unroll
      for (int k = 0; k < C / 32; ++k) {
        const float wv = wr[k * 32 + lane];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(wv, s_x[cur][r][k * 32 + lane], acc[r]);
      }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = warp_sum(acc[r]);
      if (lane < ROWS) {
        float v = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + B[l * C + j];
        if (l < L - 1) v = fmaxf(v, 0.f);
        s_x[cur ^ 1][lane][j] = v;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int r = 0; r < ROWS; ++r)
    if (r0 + r < n) dst[(long)(r0 + r) * C + tid] = s_x[cur][r][tid];
}

// One block per current row c: s[0] = 0, s[1+m] = <y_cur[c], y_bank[m]>; log_softmax over the row, max + first
// argmax (torch.max(dim=1) on CPU returns the first maximal index).  Optionally stores the score row.
__global__ void __launch_bounds__(256) track_score_kernel(const float* __restrict__ y_cur, const float* __restrict__ y_bank,
                                                          const int* __restrict__ k_ptr, int k_host,
                                                          const int* __restrict__ m_ptr, int m_host,
                                                          float* __restrict__ lik, int* __restrict__ mid,
                                                          float* __restrict__ scores, int ld_scores) {
  extern __shared__ float s_s[];                    // [1 + M]
  __shared__ float s_red[8];
  __shared__ int s_arg[8];
  const int K = k_ptr ? min(*k_ptr, k_host) : k_host;
  const int M = m_ptr ? min(*m_ptr, m_host) : m_host;
  const int c = blockIdx.x;
  if (c >= K) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float xv[C / 32];
#pragma unroll
  for (int k = 0; k < C / 32; ++k) xv[k] = y_cur[(long)c * C + k * 32 + lane];
  if (tid == 0) s_s[0] = 0.f;
  for (int m = warp; m < M; m += 8) {
    const float* br = y_bank + (long)m * C;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < C / 32; ++k) acc = fmaf(xv[k], br[k * 32 + lane], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_s[1 + m] = acc;
  }
  __syncthreads();
  float mx = -INFINITY; int am = 0x7fffffff;
  for (int j = tid; j <= M; j += 256) { const float v = s_s[j]; if (v > mx) { mx = v; am = j; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o); const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (ov > mx || (ov == mx && oa < am)) { mx = ov; am = oa; }
  }
  if (lane == 0) { s_red[warp] = mx; s_arg[warp] = am; }
  __syncthreads();
  mx = s_red[0]; am = s_arg[0];
  for (int q = 1; q < 8; ++q) if (s_red[q] > mx || (s_red[q] == mx && s_arg[q] < am)) { mx = s_red[q]; am = s_arg[q]; }
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j <= M; j += 256) sum += expf(s_s[j] - mx);
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int q = 0; q < 8; ++q) t += s_red[q];
    if (lik) lik[c] = -logf(t);                     // max of log_softmax = (mx - mx) - log(sum exp(s - mx))
    if (mid) mid[c] = am;
  }
  if (scores) for (int j = tid; j <= M; j += 256) scores[(long)c * ld_scores + j] = s_s[j];
}

// The greedy loop (:359-404).  One block; thread 0 takes the sequential decisions, then all threads update the
// bank rows.  meta = fusion meta of this frame (K' kept, keep_index[N]); out int32[4 + N]:
//   out[0] = K', out[1] = n_things, out[2] = bank rows after this frame, out[3] = overflow flag,
//   out[4 + c] = object id of kept entry c (post-processor order: stuff..., things...), -1 beyond K'.
__global__ void __launch_bounds__(256) track_assign_kernel(State* __restrict__ st, float* __restrict__ bank, int capacity,
                                                           const float* __restrict__ emb, const int* __restrict__ meta, int N,
                                                           const float* __restrict__ lik, const int* __restrict__ mid,
                                                           int* __restrict__ out) {
  extern __shared__ int s_i[];                      // best_id[capacity] | src_for_row[capacity] ; floats best[capacity]
  int* s_best_id = s_i;
  int* s_row_src = s_i + capacity;                  // bank row -> kept entry whose embedding it receives (-1: unchanged)
  float* s_best = (float*)(s_i + 2 * capacity);
  __shared__ int s_ids[1024];
  __shared__ int s_count, s_over;
  const int K = min(meta[0], N), tid = threadIdx.x;
  const int M0 = st->started ? st->count : 0;
  for (int j = tid; j < capacity; j += 256) { s_best_id[j] = -1; s_row_src[j] = -1; s_best[j] = -100.f; }
  for (int c = tid; c < N; c += 256) s_ids[c] = -1;
  __syncthreads();
  if (tid == 0) {
    int count = M0, over = 0;
    if (!st->started) {                             // first frame of the video: ids = arange(K'), bank = all entries
      for (int c = 0; c < K; ++c) {
        s_ids[c] = c;
        if (c < capacity) s_row_src[c] = c; else over = 1;
      }
      count = min(K, capacity);
      if (K > 0) st->started = 1;
    } else {
      for (int c = 0; c < K; ++c) {
        const int m = mid[c];
        if (m == 0) {                               // prefers the all-zero column: new object
          s_ids[c] = count;
          if (count < capacity) { s_row_src[count] = c; ++count; } else over = 1;
        } else {
          const int o = m - 1;
          const float l = lik[c];
          if (l > s_best[o]) {                      // best candidate so far for object o; an earlier one is undone
            s_ids[c] = o;
            if (s_best_id[o] >= 0) s_ids[s_best_id[o]] = -1;
            s_best[o] = l; s_best_id[o] = c; s_row_src[o] = c;
          }
        }
      }
      for (int c = 0; c < K; ++c) {                 // redundant matches become new objects (:396-404)
        if (s_ids[c] >= 0) continue;
        s_ids[c] = count;
        if (count < capacity) { s_row_src[count] = c; ++count; } else over = 1;
      }
    }
    s_count = count; s_over = over;
    st->count = count;
    if (over) st->overflow = 1;
  }
  __syncthreads();
  const int count = s_count;
  for (int row = tid >> 5; row < count; row += 8) { // warp per bank row
    const int c = s_row_src[row];
    if (c < 0) continue;
    const float* e = emb + (long)meta[4 + c] * C;
    for (int k = tid & 31; k < C; k += 32) bank[(long)row * C + k] = e[k];
  }
  if (tid == 0) { out[0] = K; out[1] = meta[1]; out[2] = count; out[3] = s_over | st->overflow; }
  for (int c = tid; c < N; c += 256) out[4 + c] = s_ids[c];
}

__global__ void track_reset_kernel(State* st) { st->count = 0; st->started = 0; st->overflow = 0; st->pad = 0; }

}  // namespace track
}  // namespace slotvps
