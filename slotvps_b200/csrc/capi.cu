// C ABI of libslotvps_b200.so (see include/slotvps_b200.h) and the host-side orchestration of the
// retriever hot path: level fusion -> per-stage slot update / pixel attention -> mask logits ->
// panoptic fusion.  One translation unit; kernels live in the .cuh files next to it.
#include <stdlib.h>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "sgemm.cuh"
#include "rowops.cuh"
#include "pixel_fp32.cuh"
#include "fusion.cuh"
#include "pixel_tc.cuh"
#include "fuse_tc.cuh"
#include "mask_tc.cuh"
#include "slot_tc.cuh"
#include "slot_cl.cuh"
#include "temporal.cuh"
#include "track.cuh"
#include "unify.cuh"
#include "dcn.cuh"

namespace slotvps {
thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;
thread_local bool g_prof_on = false;

struct ProfRec { const char* name; cudaEvent_t ev; int grid; };
thread_local int g_prof_grid = 0;           // CTAs of the launch about to be marked (set by the persistent-kernel launchers)
static thread_local std::vector<ProfRec> g_prof;
static thread_local std::vector<cudaEvent_t> g_prof_pool;
void prof_mark(const char* name, cudaStream_t s) {
  cudaEvent_t e;
  if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  g_prof.push_back({name, e, g_prof_grid});
  g_prof_grid = 0;
}

// epilogue activation codes of linear_fast / sgemm: 1 relu, 2 exact (erf) GELU
static int ffn_act_of(const slotvps_head_desc* d) { return d->ffn_act == 1 ? 1 : 2; }                    // default: "gelu" (r50_fpn_slotvps.py:33)
static int temporal_ffn_act_of(const slotvps_head_desc* d) { return d->temporal_ffn_act == 2 ? 2 : 1; }  // default: "relu" (r50_fpn_slotvps.py:49)

static int n_stages_of(const slotvps_head_desc* d) {
  int s = 0;
  for (int l = 0; l < d->n_levels; ++l) s += d->heads_per_level[l];
  return s;
}

static int validate(const slotvps_head_desc* d) {
  SV_REQUIRE(d != nullptr, "null descriptor");
  SV_REQUIRE(d->n_frames >= 1 && d->n_frames <= SLOTVPS_MAX_FRAMES, "n_frames out of range");
  SV_REQUIRE(d->n_slots >= 1 && d->n_slots <= 512, "n_slots out of range (1..512)");
  SV_REQUIRE(d->n_levels >= 1 && d->n_levels <= SLOTVPS_MAX_LEVELS, "n_levels out of range");
  SV_REQUIRE(n_stages_of(d) >= 1 && n_stages_of(d) <= SLOTVPS_MAX_STAGES, "stage count out of range");
  SV_REQUIRE(d->nhead * 32 == C, "nhead must be 8 (head_dim 32)");
  SV_REQUIRE(d->num_classes >= 2 && d->num_classes <= 256, "num_classes out of range");
  SV_REQUIRE(d->dim_feedforward > 0 && d->temporal_dim_feedforward > 0, "feed-forward widths");
  for (int l = 0; l < d->n_levels; ++l) {
    SV_REQUIRE(d->h[l] > 0 && d->w[l] > 0, "empty level");
    if (l > 0) SV_REQUIRE(d->h[l] == 2 * d->h[l - 1] && d->w[l] == 2 * d->w[l - 1], "each level must be 2x the previous");
  }
  SV_REQUIRE(d->pos_mode >= 0 && d->pos_mode <= 2, "pos_mode");
  SV_REQUIRE(d->ffn_act >= 0 && d->ffn_act <= 2 && d->temporal_ffn_act >= 0 && d->temporal_ffn_act <= 2, "activation code (0 default, 1 relu, 2 gelu)");
  return SLOTVPS_OK;
}

// ---- prepared (folded) weights --------------------------------------------------------------------
struct PreparedStage {
  float *Wk_c, *bk_c, *Wv_c, *bv_c;     // output-centred key/value projections
  float *Wk_cT;                         // Wk_c transposed [in][out] (G = qt . Wk_c as a K-contiguous linear)
  float *tq_qkv_w, *tq_qkv_b, *tq_ln_w, *tq_ln_b;   // Video Retriever q|k|v stacked [768,256],[768],[3,256]
  float *tw_w, *tw_ln_w, *tw_ln_b;      // first tower layers cls|reg stacked [512,256],[2,256]
  TcStageOperands tc;                   // tensor-core operand planes (pixel_tc.cuh)
  slot::SlotTcWeights stc;              // fp16 hi/lo planes of the slot-side linears (slot_tc.cuh), when slot_tc_supported
};
struct Prepared {
  float* W0;                            // level-0 folded conv weight [256,128]
  float *conv_w, *conv_b;               // conv_trans.conv.{weight [256,384], bias [256]} (with the input transform folded in)
  float* conv_b0;                       // bias of level 0 (differs from conv_b only when an input transform is folded)
  FuseTcWeights ftc;                    // fp16 hi/lo planes of the folded conv_trans weights
  PreparedStage st[SLOTVPS_MAX_STAGES];
};

static size_t prepared_layout(const slotvps_head_desc* d, void* base, Prepared* out) {
  Arena a(base, (size_t)-1);
  Prepared p;
  p.W0 = a.take<float>((size_t)C * CIN);
  p.conv_w = a.take<float>((size_t)C * 3 * CIN); p.conv_b = a.take<float>(C); p.conv_b0 = a.take<float>(C);
  p.ftc.w0 = a.take<__half>((size_t)2 * C * CIN); p.ftc.wa = a.take<__half>((size_t)2 * C * C); p.ftc.wb = a.take<__half>((size_t)2 * C * CIN);
  const int S = n_stages_of(d);
  for (int s = 0; s < S; ++s) {
    PreparedStage& ps = p.st[s];
    ps.Wk_c = a.take<float>((size_t)C * C); ps.bk_c = a.take<float>(C);
    ps.Wv_c = a.take<float>((size_t)C * C); ps.bv_c = a.take<float>(C);
    ps.Wk_cT = a.take<float>((size_t)C * C);
    ps.tq_qkv_w = a.take<float>((size_t)3 * C * C); ps.tq_qkv_b = a.take<float>(3 * C);
    ps.tq_ln_w = a.take<float>(3 * C); ps.tq_ln_b = a.take<float>(3 * C);
    ps.tw_w = a.take<float>((size_t)2 * C * C); ps.tw_ln_w = a.take<float>(2 * C); ps.tw_ln_b = a.take<float>(2 * C);
    tc_stage_layout(a, &ps.tc);
    if (slot::slot_tc_supported(d)) slot::slot_tc_layout(a, d, &ps.stc);
  }
  if (out) *out = p;
  return align_up(a.off);
}

// Wc[o][c] = W[o][c] - mean_o W[o][c] ; bc[o] = b[o] - mean(b)   (double accumulation)
__global__ void __launch_bounds__(256) center_rows_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                          float* __restrict__ Wc, float* __restrict__ bc) {
  const int c = threadIdx.x;
  double m = 0.0;
  for (int o = 0; o < C; ++o) m += (double)W[o * C + c];
  m /= C;
  for (int o = 0; o < C; ++o) Wc[o * C + c] = (float)((double)W[o * C + c] - m);
  __shared__ double sb[C];
  sb[c] = (double)b[c];
  __syncthreads();
  double mb = 0.0;
  for (int o = 0; o < C; ++o) mb += sb[o];
  bc[c] = (float)(sb[c] - mb / C);
}
__global__ void __launch_bounds__(256) transpose256_kernel(const float* __restrict__ in, float* __restrict__ out) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i < C * C) out[(i % C) * C + i / C] = in[i];
}
__global__ void __launch_bounds__(256) fold_w0_kernel(const float* __restrict__ W, float* __restrict__ W0) {
  int i = blockIdx.x * 256 + threadIdx.x;            // [256][128]
  if (i >= C * CIN) return;
  int o = i / CIN, c = i % CIN;
  const float* r = W + (long)o * (3 * CIN);
  W0[i] = (float)((double)r[c] + (double)r[CIN + c] + (double)r[2 * CIN + c]);
}
// Fold a 1x1 input transform x = T f + t (VPS_Capsule.conv_trans applied by semantic_trans_ins,
// vps_temporal_slots.py:129-135) into a conv block acting on x:  W x + b = (W T) f + (W t + b).
// One block per output row o; W rows have stride ld; in place is allowed (the row is staged first).
__global__ void __launch_bounds__(CIN) fold_in_trans_kernel(const float* __restrict__ Wsrc, int ld, const float* __restrict__ T,
                                                            const float* __restrict__ t, const float* __restrict__ b_in,
                                                            float* __restrict__ Wdst, int ld_dst, float* __restrict__ b_out) {
  __shared__ float row[CIN];
  __shared__ double red[CIN];
  const int o = blockIdx.x, c = threadIdx.x;
  row[c] = Wsrc[(long)o * ld + c];
  __syncthreads();
  double acc = 0.0;
  for (int k = 0; k < CIN; ++k) acc += (double)row[k] * (double)T[k * CIN + c];
  red[c] = (double)row[c] * (double)t[c];
  __syncthreads();
  Wdst[(long)o * ld_dst + c] = (float)acc;
  if (c == 0) {
    double sb = 0.0;
    for (int k = 0; k < CIN; ++k) sb += red[k];
    b_out[o] = (float)((double)b_in[o] + sb);
  }
}
__global__ void __launch_bounds__(256) copy_kernel(const float* __restrict__ src, float* __restrict__ dst, long n) {
  long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
static int dcopy(const float* src, float* dst, long n, cudaStream_t s) {
  copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, n);
  SV_CHECK_LAUNCH("copy");
  return SLOTVPS_OK;
}

// ---- workspace of one head invocation ----------------------------------------------------------------
struct HeadWs {
  float *slots, *qkv, *mo, *p, *qraw, *qt, *g0, *g1, *G;
  float *rs_k2, *rs_v2;                 // second LayerNorm-scale buffers (stages alternate in overlapped mode)
  float *rs_k, *rs_v, *Zpart, *a0part, *a1part, *Z, *a0, *a1, *Y, *p2, *hdn, *f, *f2;
  float *tqkv, *L, *av, *ty, *thdn, *tw, *c2, *e1;
  float *p2buf;                         // post-norm2 rows kept for the FFN residual (slot_post_kernel)
  float *ffn_part;                      // [F/128][R][256] lin2 partials of the spread FFN (slot_ffn_kernel)
  uint8_t *opx;                         // [T][106496] operand image of the slot clusters (slot_cl.cuh)
  float *pos[SLOTVPS_MAX_LEVELS];       // generated sine embeddings (pos_mode 2)
  float *ybuf;                          // coarse conv_trans partial [T][256][P/4]
  float *splitk;                        // split-K partials of the long-K linears [4][R][256]
  float *tk[SLOTVPS_MAX_STAGES];        // separable-pos key tables per stage: [h][256] then [w][256]
  float *pg[2];                         // separable-pos slot tables (stage parity): [T][h][112] then [T][w][112]
  int hmax, wmax;
  TcWorkspace tc;
  FuseTcWorkspace ftc;
  int chunks;
};
static int attn_chunks(int P, int T) {
  int tiles = ceil_div(P, ATT_PX);
  int per = 148 / (T > 0 ? T : 1);
  if (per < 1) per = 1;
  return tiles < per ? tiles : per;
}
static size_t head_ws_layout(const slotvps_head_desc* d, void* base, size_t cap, HeadWs* out) {
  Arena a(base, cap);
  HeadWs w;
  const int T = d->n_frames, N = d->n_slots, R = T * N;
  int Pmax = 0;
  for (int l = 0; l < d->n_levels; ++l) Pmax = max(Pmax, d->h[l] * d->w[l]);
  const int chunks = 148;               // upper bound of attn_chunks * NB-independent
  w.chunks = chunks;
  w.slots = a.take<float>((size_t)R * C); w.qkv = a.take<float>((size_t)R * 3 * C); w.mo = a.take<float>((size_t)R * C);
  w.p = a.take<float>((size_t)R * C); w.qraw = a.take<float>((size_t)R * C); w.qt = a.take<float>((size_t)R * C);
  w.g0 = a.take<float>(R); w.g1 = a.take<float>(R); w.G = a.take<float>((size_t)R * C);
  w.rs_k = a.take<float>((size_t)T * Pmax); w.rs_v = a.take<float>((size_t)T * Pmax);
  w.rs_k2 = a.take<float>((size_t)T * Pmax); w.rs_v2 = a.take<float>((size_t)T * Pmax);
  w.Zpart = a.take<float>((size_t)chunks * R * C); w.a0part = a.take<float>((size_t)chunks * R); w.a1part = a.take<float>((size_t)chunks * R);
  w.Z = a.take<float>((size_t)R * C); w.a0 = a.take<float>(R); w.a1 = a.take<float>(R);
  w.Y = a.take<float>((size_t)R * C); w.p2 = a.take<float>((size_t)R * C);
  w.hdn = a.take<float>((size_t)R * max(d->dim_feedforward, d->temporal_dim_feedforward));
  w.f = a.take<float>((size_t)R * C); w.f2 = a.take<float>((size_t)R * C);
  w.tqkv = a.take<float>((size_t)R * 3 * C); w.L = a.take<float>((size_t)R * R); w.av = a.take<float>((size_t)R * C);
  w.ty = a.take<float>((size_t)R * C); w.thdn = w.hdn;
  w.tw = a.take<float>((size_t)R * 2 * C); w.c2 = a.take<float>((size_t)R * C); w.e1 = a.take<float>((size_t)R * C);
  w.p2buf = a.take<float>((size_t)R * C);
  w.ffn_part = a.take<float>((size_t)(max(d->dim_feedforward, d->temporal_dim_feedforward) / 128 + 1) * R * C);
  w.opx = a.take<uint8_t>((size_t)T * slot::slot_groups(d->n_slots) * slot::ACT_BYTES);
  for (int l = 0; l < SLOTVPS_MAX_LEVELS; ++l)
    w.pos[l] = (d->pos_mode == 2 && l < d->n_levels) ? a.take<float>((size_t)C * d->h[l] * d->w[l]) : nullptr;
  w.ybuf = a.take<float>((size_t)T * C * (Pmax / 4 + 1));
  w.splitk = a.take<float>((size_t)4 * R * C);
  {
    int hm = 0, wm = 0;
    for (int l = 0; l < d->n_levels; ++l) { hm = max(hm, d->h[l]); wm = max(wm, d->w[l]); }
    w.hmax = hm; w.wmax = wm;
    for (int i = 0; i < SLOTVPS_MAX_STAGES; ++i) w.tk[i] = (d->kernel_path == 0) ? a.take<float>((size_t)(hm + wm) * C) : nullptr;
    for (int i = 0; i < 2; ++i) w.pg[i] = (d->kernel_path == 0) ? a.take<float>((size_t)T * (hm + wm) * 112) : nullptr;
  }
  tc_workspace_layout(a, d, &w.tc);
  if (d->kernel_path == 0) {
    w.ftc.in_planes = a.take<__half>((size_t)2 * T * Pmax * CIN);
    w.ftc.y = a.take<float>((size_t)T * (Pmax / 4 + 1) * C);
  }
  if (out) *out = w;
  return align_up(a.off);
}

// ---- two-stream schedule ---------------------------------------------------------------------------------
// The slot-side chain of a stage is ~30 short dependent kernels on few SMs; the pixel-side producers
// (level fusion, LayerNorm statistics) only depend on features.  In overlapped mode they run on an
// internal side stream with a reduced grid, the attention kernel joins both (events), so the tensor-pipe
// work hides behind the latency-bound slot chain.  Captured CUDA graphs keep the fork/join structure.
struct Overlap {
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> ev;
  int next = 0;
  int init() {
    if (!side) SV_CHECK_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    while (ev.size() < 64) {
      cudaEvent_t e;
      SV_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev.push_back(e);
    }
    next = 0;
    return SLOTVPS_OK;
  }
  cudaEvent_t take() { return ev[next++ % ev.size()]; }
};
// A small pool of side streams handed out round-robin, created on the first (never captured: callers warm up
// before capturing) call, so concurrent clips on different caller streams do not serialise on one side stream.
static thread_local Overlap g_ovs[4];
static thread_local unsigned g_ov_next = 0;
static Overlap& overlap_next() { return g_ovs[g_ov_next++ % 4]; }
static thread_local bool g_last_head_alt = false;   // did the last head_forward leave the finest level in planes_alt?
constexpr int SIDE_CTAS = 112;           // tensor-core producers leave 36 SMs to the slot-side kernels

struct StagePix {                        // how a stage gets its pixel-side inputs
  bool overlapped = false;               // true: rs_k/rs_v are produced on the side stream, wait for `ready`
  cudaEvent_t ready = nullptr, done = nullptr;
  float *rs_k = nullptr, *rs_v = nullptr;
  TcWorkspace tc;
  PosSep ps;                             // separable-pos tables (enabled when the level's planes carry x only)
  float* pg = nullptr;                   // where this stage's pgy|pgx tables go
};

// ---- level fusion (dynamic_mask_head.py:172-185), folded: W.cat(up(p),x) = up(Wa.p) + Wb.x ----------
static int level_fuse_frame(const float* prev, const float* x, const float* conv_w, const float* conv_b, const float* W0,
                            float* out, int h, int w, float* ybuf, cudaStream_t s) {
  const int P = h * w;
  GemmArgs g;
  g.B = x; g.b_ks = P; g.b_ns = 1;
  g.Cm = out; g.c_ms = P; g.c_ns = 1;
  g.M = C; g.N = P; g.K = CIN;
  g.bias = conv_b; g.bias_mode = 1;
  if (!prev) {
    g.A = W0; g.a_ms = CIN; g.a_ks = 1;
    return sgemm(g, s);
  }
  const int Pc = (h / 2) * (w / 2);
  GemmArgs gy;
  gy.A = conv_w; gy.a_ms = 3 * CIN; gy.a_ks = 1;
  gy.B = prev; gy.b_ks = Pc; gy.b_ns = 1;
  gy.Cm = ybuf; gy.c_ms = Pc; gy.c_ns = 1;
  gy.M = C; gy.N = Pc; gy.K = C;
  SV_TRY(sgemm(gy, s));
  g.A = conv_w + 2 * CIN; g.a_ms = 3 * CIN; g.a_ks = 1;
  g.up = ybuf; g.up_h = h / 2; g.up_w = w / 2;
  return sgemm(g, s);
}

// ---- pixel attention, fp32 path: rs_k, rs_v, then fused S/softmax/Z, then deterministic reduce --------
// (use_tc: the LayerNorm statistics -- 72 % of the contraction's FLOPs -- run on the tensor pipe from the
//  level's bf16 operand planes, pixel_tc.cuh; the slot-softmax contraction below still runs in fp32)
static int pixel_attention(const float* x, long x_bs, const float* pos, long pos_bs, const PreparedStage& ps,
                           const HeadWs& w0, const StagePix& px, int T, int N, int P, bool use_tc, cudaStream_t s, bool gplanes_ready = false) {
  HeadWs w = w0;
  w.rs_k = px.rs_k; w.rs_v = px.rs_v; w.tc = px.tc;
  if (px.overlapped) {
    SV_CHECK_CUDA(cudaStreamWaitEvent(s, px.ready, 0));      // statistics of this stage were produced on the side stream
  } else if (use_tc && !(x != nullptr && getenv("SLOTVPS_TC_DEBUG") && (atoi(getenv("SLOTVPS_TC_DEBUG")) & 64))) {
    SV_TRY(tc_stats(ps.tc, w.tc, ps.bk_c, ps.bv_c, w.rs_k, w.rs_v, T, P, s, 148, px.ps));
  } else {
    dim3 gs(ceil_div(P, 32), T);
    proj_rstd_kernel<<<gs, 256, 0, s>>>(x, x_bs, pos, pos_bs, ps.Wk_c, ps.bk_c, w.rs_k, P);
    SV_CHECK_LAUNCH("proj_rstd(k)");
    proj_rstd_kernel<<<gs, 256, 0, s>>>(x, x_bs, nullptr, 0, ps.Wv_c, ps.bv_c, w.rs_v, P);
    SV_CHECK_LAUNCH("proj_rstd(v)");
  }
  int chunks = 0;
  if (use_tc) {
    PosSep pa = px.ps;
    if (pa.enabled && pa.tky != nullptr) {
      // pgy[t][row][n] = sum_{c<128} ytab[c][row] G[t][n][c] ; pgx[t][col][n] = sum_{c<128} xtab[c][col] G[t][n][128+c]
      float* pgy = px.pg;
      float* pgx = px.pg + (long)T * pa.h * attn::NPAD;
      GemmArgs g;
      g.A = w.tc.ytab; g.a_ms = 1; g.a_ks = pa.h;
      g.B = w.G; g.b_ks = 1; g.b_ns = C; g.b_bs = (long)N * C;
      g.Cm = pgy; g.c_ms = attn::NPAD; g.c_ns = 1; g.c_bs = (long)pa.h * attn::NPAD;
      g.M = pa.h; g.N = N; g.K = 128; g.batch = T;
      SV_TRY(sgemm(g, s));
      g.A = w.tc.xtab; g.a_ks = pa.w;
      g.B = w.G + 128;
      g.Cm = pgx; g.c_bs = (long)pa.w * attn::NPAD;
      g.M = pa.w;
      SV_TRY(sgemm(g, s));
      pa.pgy = pgy; pa.pgx = pgx;
    }
    SV_TRY(tc_attention(w.tc, w.tc.gplanes, w.G, w.g0, w.g1, w.rs_k, w.rs_v, w.Zpart, w.a0part, w.a1part, T, N, P, &chunks, s, pa, gplanes_ready));
  } else {
    const int NB = ceil_div(N, 128);
    chunks = attn_chunks(P, T);
    dim3 ga(chunks, T, NB);
    const size_t smem = att_fp32_smem_bytes();
#define SV_ATT(NBV)                                                                                                  \
    {                                                                                                                  \
      SV_TRY(ensure_dyn_smem((const void*)slot_attn_fp32_kernel<NBV>, smem));                                         \
      slot_attn_fp32_kernel<NBV><<<ga, 256, smem, s>>>(x, x_bs, pos, pos_bs, w.G, w.g0, w.g1, w.rs_k, w.rs_v, w.Zpart,  \
                                                       w.a0part, w.a1part, N, P, T);                                   \
    }
    switch (NB) {
      case 1: SV_ATT(1) break;
      case 2: SV_ATT(2) break;
      case 3: SV_ATT(3) break;
      case 4: SV_ATT(4) break;
      default: return fail(SLOTVPS_EINVAL, "n_slots > 512 unsupported%s%s");
    }
#undef SV_ATT
    SV_CHECK_LAUNCH("slot_attn_fp32");
  }
  if (px.overlapped) SV_CHECK_CUDA(cudaEventRecord(px.done, s));   // planes / rs buffers of this stage may be recycled
  const long nz = (long)T * N * C, na = (long)T * N;
  reduce_attn_parts_kernel<<<(unsigned)((nz / 4 + 2 * na + 31) / 32), 256, 0, s>>>(w.Zpart, w.a0part, w.a1part, w.Z, w.a0, w.a1, nz, na, chunks);
  SV_CHECK_LAUNCH("reduce_attn_parts");
  return SLOTVPS_OK;
}

// ---- Video Retriever over the T*N slots of all frames (:308-322, 494-527, 550-572): f -> f2 = f + block(f) --------
static int video_retriever(const slotvps_head_desc* d, const slotvps_stage_params& sp, const PreparedStage& ps, const HeadWs& w, cudaStream_t s) {
  const int T = d->n_frames, N = d->n_slots, R = T * N, TF = d->temporal_dim_feedforward;
  SV_REQUIRE(sp.tq_to_q_w != nullptr, "temporal stage without temporal_query_head parameters");
  SV_TRY(linear_fast(w.f, ps.tq_qkv_w, ps.tq_qkv_b, w.tqkv, R, C, 3 * C, 0, nullptr, s));
  SV_TRY(ln_rows(w.tqkv, nullptr, ps.tq_ln_w, ps.tq_ln_b, 3, nullptr, w.tqkv, 3 * R, 0, s));   // rows r*3+{q,k,v}
  {
    GemmArgs g;                                       // L[l,u] = q_l . k_u
    g.A = w.tqkv; g.a_ms = 3 * C; g.a_ks = 1;
    g.B = w.tqkv + C; g.b_ks = 1; g.b_ns = 3 * C;
    g.Cm = w.L; g.c_ms = R; g.c_ns = 1;
    g.M = R; g.N = R; g.K = C;
    SV_TRY(sgemm(g, s));
  }
  col_softmax_kernel<<<ceil_div(R, 32), 256, 0, s>>>(w.L, R);
  SV_CHECK_LAUNCH("col_softmax");
  {
    GemmArgs g;                                       // av = A . v
    g.A = w.L; g.a_ms = R; g.a_ks = 1;
    g.B = w.tqkv + 2 * C; g.b_ks = 3 * C; g.b_ns = 1;
    g.Cm = w.av; g.c_ms = C; g.c_ns = 1;
    g.M = R; g.N = C; g.K = R;
    SV_TRY(sgemm(g, s));
  }
  SV_TRY(ln_rows(w.av, nullptr, sp.tq_no_w, sp.tq_no_b, 1, w.f, w.ty, R, 1, s));              // f + relu(LN(av))
  SV_TRY(ln_rows(w.ty, nullptr, sp.tq_norm2_w, sp.tq_norm2_b, 1, nullptr, w.ty, R, 0, s));    // y
  SV_TRY(linear_fast(w.ty, sp.tq_lin1_w, sp.tq_lin1_b, w.thdn, R, C, TF, temporal_ffn_act_of(d), nullptr, s));
  SV_TRY(linear_fast(w.thdn, sp.tq_lin2_w, sp.tq_lin2_b, w.qraw, R, TF, C, 0, w.ty, s, -1, -1, w.splitk));
  SV_TRY(ln_rows(w.qraw, nullptr, sp.tq_norm3_w, sp.tq_norm3_b, 1, w.f, w.f2, R, 0, s));      // X + LN3(...)  (:317)
  return SLOTVPS_OK;
}

// The same block on the slot kernels: q|k|v projections + LayerNorms on the frame clusters (slot_cl.cuh), the R x R attention
// and its two LayerNorms in two fp32 launches (temporal.cuh), the FFN over (frame, hidden chunk) CTAs + reduction/norm3.
static int video_retriever_tc(const slotvps_head_desc* d, const slotvps_stage_params& sp, const PreparedStage& ps, const HeadWs& w, cudaStream_t s) {
  const int T = d->n_frames, N = d->n_slots, R = T * N, TF = d->temporal_dim_feedforward, G = slot::slot_groups(N);
  SV_REQUIRE(sp.tq_to_q_w != nullptr, "temporal stage without temporal_query_head parameters");
  {
    CUtensorMap m_qkv;
    SV_TRY(slot::cl::slot_wmap64(&m_qkv, ps.stc.tqkv, 3 * C, C));
    slot::cl::TqkvParams tp;
    tp.N = N; tp.G = G; tp.f_in = w.f; tp.bias = ps.tq_qkv_b; tp.ln_w = ps.tq_ln_w; tp.ln_b = ps.tq_ln_b; tp.tqkv = w.tqkv;
    SV_TRY(ensure_dyn_smem((const void*)slot::cl::slot_tqkv_cl, slot::cl::SMEM));
    slot::cl::slot_tqkv_cl<<<T * G * slot::cl::CL, slot::THREADS, slot::cl::SMEM, s>>>(m_qkv, tp);
    SV_CHECK_LAUNCH("slot_tqkv");
  }
  temporal::tscore_kernel<<<ceil_div(R, temporal::KEYS), temporal::TS_WARPS * 32, 0, s>>>(w.tqkv, w.L, R);
  SV_CHECK_LAUNCH("tscore");
  SV_TRY(ensure_dyn_smem((const void*)temporal::tav_kernel, temporal::TAV_SMEM));
  temporal::tav_kernel<<<ceil_div(R, temporal::TAV_ROWS), temporal::TAV_ROWS * 32, temporal::TAV_SMEM, s>>>(
      w.L, w.tqkv, w.f, sp.tq_no_w, sp.tq_no_b, sp.tq_norm2_w, sp.tq_norm2_b, w.ty, R);
  SV_CHECK_LAUNCH("tav");
  {
    CUtensorMap m_l1, m_l2;
    SV_TRY(slot::slot_wmap(&m_l1, ps.stc.tlin1, TF, C));
    SV_TRY(slot::slot_wmap(&m_l2, ps.stc.tlin2, C, TF));
    slot::FfnParams fp;
    fp.N = N; fp.G = G; fp.act = temporal_ffn_act_of(d); fp.p2 = w.ty; fp.b1 = sp.tq_lin1_b; fp.part = w.ffn_part; fp.part_stride = (long)R * C;
    SV_TRY(ensure_dyn_smem((const void*)slot::slot_ffn_kernel, slot::SMEM_BYTES));
    slot::slot_ffn_kernel<<<dim3(T * G, TF / 128), slot::THREADS, slot::SMEM_BYTES, s>>>(m_l1, m_l2, fp, TF);
    SV_CHECK_LAUNCH("slot_ffn");
    slot::slot_norm3_kernel<<<ceil_div(R, 8), 256, 0, s>>>(w.ffn_part, (long)R * C, TF / 128, w.ty, sp.tq_lin2_b, sp.tq_norm3_w, sp.tq_norm3_b,
                                                           w.f, w.f2, R);                        // X + LN3(...)  (:317)
    SV_CHECK_LAUNCH("slot_norm3");
  }
  return SLOTVPS_OK;
}

static int pixel_attention(const float* x, long x_bs, const float* pos, long pos_bs, const PreparedStage& ps,
                           const HeadWs& w0, const StagePix& px, int T, int N, int P, bool use_tc, cudaStream_t s, bool gplanes_ready);

// ---- the same stage with the slot side on the tensor-core slot-update kernels (slot_tc.cuh): N <= 104 ------------
static int run_stage_tc(const slotvps_head_desc* d, const slotvps_stage_params& sp, const PreparedStage& ps, const HeadWs& w,
                        const StagePix& px, const float* x, long x_bs, const float* pos, long pos_bs, int h, int wd, bool temporal,
                        float* cls_out, long cls_frame_stride, float* emb_out, long emb_frame_stride, cudaStream_t s) {
  const int T = d->n_frames, N = d->n_slots, R = T * N, P = h * wd, F = d->dim_feedforward;
  const int G = slot::slot_groups(N);                     // row groups of <= 104 slots per frame
  const int grid_cl = T * G * slot::cl::CL;               // one cluster of four CTAs per (frame, group) (slot_cl.cuh)
  // (1) slot self-attention core (:346-352): in_proj + 8-head attention on the generic kernels
  SV_TRY(linear_fast(w.slots, sp.in_proj_w, sp.in_proj_b, w.qkv, R, C, 3 * C, 0, nullptr, s));
  {
    size_t smem = (size_t)(2 * N * 33 + 8 * N) * sizeof(float);
    if (smem > 48 * 1024) SV_TRY(ensure_dyn_smem((const void*)mha_core_kernel, smem));
    mha_core_kernel<<<dim3(d->nhead, T, 4), 256, smem, s>>>(w.qkv, w.mo, N, d->nhead);
    SV_CHECK_LAUNCH("mha_core");
  }
  // (2) out_proj + norm1, to_q + norm_q, folded key operands G / g0 / g1 and their fp16 planes
  {
    CUtensorMap m_out, m_q, m_wk;
    SV_TRY(slot::cl::slot_wmap64(&m_out, ps.stc.out_proj, C, C));
    SV_TRY(slot::cl::slot_wmap64(&m_q, ps.stc.to_q, C, C));
    SV_TRY(slot::cl::slot_wmap64(&m_wk, ps.stc.wkT, C, C));
    slot::PreParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.N = N; pp.G = G; pp.mo = w.mo; pp.slots = w.slots;
    pp.dbg = getenv("SLOTVPS_SLOT_DEBUG") ? atoi(getenv("SLOTVPS_SLOT_DEBUG")) : 0;
    pp.out_b = sp.out_proj_b; pp.n1_w = sp.norm1_w; pp.n1_b = sp.norm1_b; pp.q_b = sp.to_q_b;
    pp.nq_w = sp.nq_w; pp.nq_b = sp.nq_b; pp.nk_w = sp.nk_w; pp.nk_b = sp.nk_b; pp.bk_c = ps.bk_c;
    pp.p = w.p; pp.Gout = w.G; pp.g0 = w.g0; pp.g1 = w.g1; pp.gplanes = G == 1 ? px.tc.gplanes : nullptr; pp.opx = w.opx;
    SV_TRY(ensure_dyn_smem((const void*)slot::cl::slot_pre_cl, slot::cl::SMEM));
    slot::cl::slot_pre_cl<<<grid_cl, slot::THREADS, slot::cl::SMEM, s>>>(m_out, m_q, m_wk, pp);
    SV_CHECK_LAUNCH("slot_pre");
  }
  // (3) pixel side: Z, a0, a1
  SV_TRY(pixel_attention(x, x_bs, pos, pos_bs, ps, w, px, T, N, P, true, s, G == 1));
  // (4..7) value projection + norms, FFN, norm3 [, Video Retriever], towers
  {
    CUtensorMap m_wv, m_l1, m_l2, m_tw, m_c1, m_r1, m_lg;
    SV_TRY(slot::cl::slot_wmap64(&m_wv, ps.stc.wv, C, C));
    SV_TRY(slot::slot_wmap(&m_l1, ps.stc.lin1, F, C));
    SV_TRY(slot::slot_wmap(&m_l2, ps.stc.lin2, C, F));
    SV_TRY(slot::cl::slot_wmap64(&m_tw, ps.stc.tw, 2 * C, C));
    SV_TRY(slot::cl::slot_wmap64(&m_c1, ps.stc.cls1, C, C));
    SV_TRY(slot::cl::slot_wmap64(&m_r1, ps.stc.reg1, C, C));
    SV_TRY(slot::cl::slot_wmap64(&m_lg, ps.stc.logit, d->num_classes, C));
    slot::PostParams q;
    memset(&q, 0, sizeof(q));
    q.N = N; q.G = G; q.ncls = d->num_classes;
    q.Z = w.Z; q.a0 = w.a0; q.a1 = w.a1; q.p = w.p;
    q.nv_w = sp.nv_w; q.nv_b = sp.nv_b; q.bv_c = ps.bv_c; q.no_w = sp.no_w; q.no_b = sp.no_b; q.n2_w = sp.norm2_w; q.n2_b = sp.norm2_b;
    q.p2buf = w.p2buf; q.opx = w.opx;
    q.tw_ln_w = ps.tw_ln_w; q.tw_ln_b = ps.tw_ln_b; q.c1_nw = sp.cls1_nw; q.c1_nb = sp.cls1_nb; q.r1_nw = sp.reg1_nw; q.r1_nb = sp.reg1_nb;
    q.logit_b = sp.logit_b;
    q.slots_out = w.slots; q.emb_out = emb_out; q.cls_out = cls_out; q.emb_fs = emb_frame_stride; q.cls_fs = cls_frame_stride;
    SV_TRY(ensure_dyn_smem((const void*)slot::cl::slot_post_cl, slot::cl::SMEM));
    slot::cl::slot_post_cl<<<grid_cl, slot::THREADS, slot::cl::SMEM, s>>>(m_wv, q);
    SV_CHECK_LAUNCH("slot_post");
    // FFN over (frame, hidden chunk) CTAs, then the chunk-ordered reduction + residual + norm3
    slot::FfnParams fp;
    fp.N = N; fp.G = G; fp.act = ffn_act_of(d); fp.p2 = w.p2buf; fp.b1 = sp.lin1_b; fp.part = w.ffn_part; fp.part_stride = (long)R * C;
    SV_TRY(ensure_dyn_smem((const void*)slot::slot_ffn_kernel, slot::SMEM_BYTES));
    slot::slot_ffn_kernel<<<dim3(T * G, F / 128), slot::THREADS, slot::SMEM_BYTES, s>>>(m_l1, m_l2, fp, F);
    SV_CHECK_LAUNCH("slot_ffn");
    slot::slot_norm3_kernel<<<ceil_div(R, 8), 256, 0, s>>>(w.ffn_part, (long)R * C, F / 128, w.p2buf, sp.lin2_b, sp.norm3_w, sp.norm3_b, nullptr, w.f, R);
    SV_CHECK_LAUNCH("slot_norm3");
    const float* fcur = w.f;
    if (temporal) {
      if (R <= temporal::RMAX) SV_TRY(video_retriever_tc(d, sp, ps, w, s));
      else SV_TRY(video_retriever(d, sp, ps, w, s));
      fcur = w.f2;
    }
    q.f_in = fcur;
    SV_TRY(ensure_dyn_smem((const void*)slot::cl::slot_towers_cl, slot::cl::SMEM));
    slot::cl::slot_towers_cl<<<grid_cl, slot::THREADS, slot::cl::SMEM, s>>>(m_tw, m_c1, m_lg, m_r1, q);
    SV_CHECK_LAUNCH("slot_towers");
  }
  return SLOTVPS_OK;
}

// ---- one MaskRCNNHead stage for all frames (dynamic_mask_head.py:291-400) ------------------------------
static int run_stage(const slotvps_head_desc* d, const slotvps_stage_params& sp, const PreparedStage& ps, const HeadWs& w,
                     const StagePix& px, const float* x, long x_bs, const float* pos, long pos_bs, int h, int wd, bool temporal, bool use_tc,
                     float* cls_out /*[T][S][N][K] base at this stage*/, long cls_frame_stride,
                     float* emb_out, long emb_frame_stride, const float* const* slots_in /*[T] or null*/, cudaStream_t s) {
  const int T = d->n_frames, N = d->n_slots, R = T * N, P = h * wd, F = d->dim_feedforward;
  if (slots_in)                                              // teacher forcing: this stage's slots come from the caller
    for (int t = 0; t < T; ++t)
      if (slots_in[t]) SV_TRY(dcopy(slots_in[t], w.slots + (long)t * N * C, (long)N * C, s));
  static const int slot_tc_on = getenv("SLOTVPS_SLOT_TC") ? atoi(getenv("SLOTVPS_SLOT_TC")) : 1;
  if (slot_tc_on && use_tc && slot::slot_tc_supported(d))
    return run_stage_tc(d, sp, ps, w, px, x, x_bs, pos, pos_bs, h, wd, temporal, cls_out, cls_frame_stride, emb_out, emb_frame_stride, s);
  // (1) slot self-attention + norm1  (:346-358)
  SV_TRY(linear_fast(w.slots, sp.in_proj_w, sp.in_proj_b, w.qkv, R, C, 3 * C, 0, nullptr, s));
  {
    size_t smem = (size_t)(2 * N * 33 + 8 * N) * sizeof(float);
    if (smem > 48 * 1024) SV_TRY(ensure_dyn_smem((const void*)mha_core_kernel, smem));
    mha_core_kernel<<<dim3(d->nhead, T, 4), 256, smem, s>>>(w.qkv, w.mo, N, d->nhead);
    SV_CHECK_LAUNCH("mha_core");
  }
  SV_TRY(linear_fast(w.mo, sp.out_proj_w, sp.out_proj_b, w.qraw, R, C, C, 0, w.slots, s));       // s + attn
  SV_TRY(ln_rows(w.qraw, nullptr, sp.norm1_w, sp.norm1_b, 1, nullptr, w.p, R, 0, s));       // p
  // (2) query side of the Panoptic Retriever (:431) + folded key operands
  SV_TRY(linear_fast(w.p, sp.to_q_w, sp.to_q_b, w.qraw, R, C, C, 0, nullptr, s));
  q_post_kernel<<<ceil_div(R, 8), 256, 0, s>>>(w.qraw, sp.nq_w, sp.nq_b, sp.nk_w, sp.nk_b, ps.bk_c, w.qt, w.g0, w.g1, R);
  SV_CHECK_LAUNCH("q_post");
  SV_TRY(linear_fast(w.qt, ps.Wk_cT, nullptr, w.G, R, C, C, 0, nullptr, s));     // G = qt . Wk_c  (Wk_cT = Wk_c^T, [c][o])
  // (3) pixel side: Z, a0, a1
  SV_TRY(pixel_attention(x, x_bs, pos, pos_bs, ps, w, px, T, N, P, use_tc, s));
  // (4) value projection on the pixel-reduced slots, norm1/ReLU, residual, norm2 (:456-459, 374-376)
  SV_TRY(linear_fast(w.Z, ps.Wv_c, nullptr, w.Y, R, C, C, 0, nullptr, s));
  attn_post_kernel<<<ceil_div(R, 8), 256, 0, s>>>(w.Y, w.a0, w.a1, w.p, sp.nv_w, sp.nv_b, ps.bv_c, sp.no_w, sp.no_b,
                                                   sp.norm2_w, sp.norm2_b, nullptr, w.p2, R);
  SV_CHECK_LAUNCH("attn_post");
  // (5) FFN + norm3 (:379-385); activation = exact GELU (r50 config) or ReLU (swinL config)
  SV_TRY(linear_fast(w.p2, sp.lin1_w, sp.lin1_b, w.hdn, R, C, F, ffn_act_of(d), nullptr, s));
  SV_TRY(linear_fast(w.hdn, sp.lin2_w, sp.lin2_b, w.qraw, R, F, C, 0, w.p2, s, -1, -1, w.splitk));
  SV_TRY(ln_rows(w.qraw, nullptr, sp.norm3_w, sp.norm3_b, 1, nullptr, w.f, R, 0, s));
  // (6) Video Retriever over the T*N slots of all frames (:308-322, 494-527, 550-572)
  const float* fcur = w.f;
  if (temporal) {
    SV_TRY(video_retriever(d, sp, ps, w, s));
    fcur = w.f2;
  }
  // (7) towers (:390-400): first layers of cls|reg share the input
  SV_TRY(linear_fast(fcur, ps.tw_w, nullptr, w.tw, R, C, 2 * C, 0, nullptr, s));
  SV_TRY(ln_rows(w.tw, nullptr, ps.tw_ln_w, ps.tw_ln_b, 2, nullptr, w.tw, 2 * R, 1, s));         // rows r*2+{cls,reg}
  SV_TRY(linear_fast(w.tw, sp.cls1_w, nullptr, w.c2, R, C, C, 0, nullptr, s, 2 * C));
  SV_TRY(ln_rows(w.c2, nullptr, sp.cls1_nw, sp.cls1_nb, 1, nullptr, w.c2, R, 1, s));
  SV_TRY(linear_fast(w.tw + C, sp.reg1_w, nullptr, w.e1, R, C, C, 0, nullptr, s, 2 * C));
  SV_TRY(ln_rows(w.e1, nullptr, sp.reg1_nw, sp.reg1_nb, 1, nullptr, w.slots, R, 1, s));          // next-stage slots
  for (int t = 0; t < T; ++t) {
    SV_TRY(linear_fast(w.c2 + (long)t * N * C, sp.logit_w, sp.logit_b, cls_out + t * cls_frame_stride, N, C, d->num_classes, 0, nullptr, s));
    SV_TRY(dcopy(w.slots + (long)t * N * C, emb_out + t * emb_frame_stride, (long)N * C, s));
  }
  return SLOTVPS_OK;
}

}  // namespace slotvps

using namespace slotvps;

extern "C" {

const char* slotvps_last_error(void) { return g_err; }
const char* slotvps_version(void) { return "slotvps_b200 0.1 (sm_100a)"; }
int64_t slotvps_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// ---- per-launch profiling (see common.cuh) ----------------------------------------------------------
int slotvps_profile_begin(void* stream) {
  for (auto& r : g_prof) g_prof_pool.push_back(r.ev);
  g_prof.clear();
  g_prof_on = true;
  prof_mark("(begin)", (cudaStream_t)stream);
  return SLOTVPS_OK;
}
// Synchronises the recorded events; writes up to `cap` rows "name\tlaunches\ttotal_ms\n" into buf.
int slotvps_profile_end(char* buf, size_t cap) {
  g_prof_on = false;
  SV_REQUIRE(buf && cap > 0, "null buffer");
  buf[0] = 0;
  if (g_prof.empty()) return SLOTVPS_OK;
  SV_CHECK_CUDA(cudaEventSynchronize(g_prof.back().ev));
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<int> cnt, grid;
  for (size_t i = 1; i < g_prof.size(); ++i) {
    float t = 0.f;
    SV_CHECK_CUDA(cudaEventElapsedTime(&t, g_prof[i - 1].ev, g_prof[i].ev));
    size_t k = 0;
    for (; k < names.size(); ++k) if (names[k] == g_prof[i].name) break;
    if (k == names.size()) { names.push_back(g_prof[i].name); ms.push_back(0.0); cnt.push_back(0); grid.push_back(0); }
    ms[k] += t; cnt[k] += 1;
    if (g_prof[i].grid > grid[k]) grid[k] = g_prof[i].grid;
  }
  size_t off = 0;
  for (size_t k = 0; k < names.size(); ++k) {
    int n = snprintf(buf + off, cap - off, "%s\t%d\t%.6f\t%d\n", names[k].c_str(), cnt[k], ms[k], grid[k]);
    if (n < 0 || (size_t)n >= cap - off) break;
    off += n;
  }
  return SLOTVPS_OK;
}

int slotvps_head_workspace_bytes(const slotvps_head_desc* d, size_t* bytes) {
  SV_TRY(validate(d));
  SV_REQUIRE(bytes != nullptr, "null out pointer");
  *bytes = head_ws_layout(d, nullptr, (size_t)-1, nullptr);
  return SLOTVPS_OK;
}

int slotvps_prepared_bytes(const slotvps_head_desc* d, size_t* bytes) {
  SV_TRY(validate(d));
  SV_REQUIRE(bytes != nullptr, "null out pointer");
  *bytes = prepared_layout(d, nullptr, nullptr);
  return SLOTVPS_OK;
}

int slotvps_prepare_weights(const slotvps_head_desc* d, const slotvps_stage_params* stages, const float* conv_w,
                            const float* conv_b, void* prepared, void* stream) {
  return slotvps_prepare_weights_ex(d, stages, conv_w, conv_b, nullptr, nullptr, prepared, stream);
}

int slotvps_prepare_weights_ex(const slotvps_head_desc* d, const slotvps_stage_params* stages, const float* conv_w,
                               const float* conv_b, const float* in_trans_w, const float* in_trans_b, void* prepared,
                               void* stream) {
  SV_TRY(validate(d));
  SV_REQUIRE(stages && conv_w && conv_b && prepared, "null argument");
  SV_REQUIRE((in_trans_w == nullptr) == (in_trans_b == nullptr), "in_trans_w and in_trans_b go together");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  Prepared p;
  prepared_layout(d, prepared, &p);
  fold_w0_kernel<<<ceil_div(C * CIN, 256), 256, 0, s>>>(conv_w, p.W0);
  SV_CHECK_LAUNCH("fold_w0");
  SV_TRY(dcopy(conv_w, p.conv_w, (long)C * 3 * CIN, s));
  SV_TRY(dcopy(conv_b, p.conv_b, C, s));
  SV_TRY(dcopy(conv_b, p.conv_b0, C, s));
  if (in_trans_w) {
    // the head then takes the UN-transformed semantic-head features: level 0 (W0 = Wa1 + Wa2 + Wb on cat(x,x,x)) and
    // the x block of the other levels absorb T; each gets its own bias
    fold_in_trans_kernel<<<C, CIN, 0, s>>>(p.W0, CIN, in_trans_w, in_trans_b, conv_b, p.W0, CIN, p.conv_b0);
    SV_CHECK_LAUNCH("fold_in_trans");
    fold_in_trans_kernel<<<C, CIN, 0, s>>>(p.conv_w + 2 * CIN, 3 * CIN, in_trans_w, in_trans_b, conv_b, p.conv_w + 2 * CIN, 3 * CIN, p.conv_b);
    SV_CHECK_LAUNCH("fold_in_trans");
  }
  conv_planes_kernel<<<ceil_div(C * CIN, 256), 256, 0, s>>>(p.W0, CIN, 0, CIN, p.ftc.w0);
  SV_CHECK_LAUNCH("conv_planes");
  conv_planes_kernel<<<ceil_div(C * C, 256), 256, 0, s>>>(p.conv_w, 3 * CIN, 0, C, p.ftc.wa);
  SV_CHECK_LAUNCH("conv_planes");
  conv_planes_kernel<<<ceil_div(C * CIN, 256), 256, 0, s>>>(p.conv_w, 3 * CIN, 2 * CIN, CIN, p.ftc.wb);
  SV_CHECK_LAUNCH("conv_planes");
  const int S = n_stages_of(d);
  for (int i = 0; i < S; ++i) {
    const slotvps_stage_params& sp = stages[i];
    PreparedStage& ps = p.st[i];
    center_rows_kernel<<<1, 256, 0, s>>>(sp.to_k_w, sp.to_k_b, ps.Wk_c, ps.bk_c);
    SV_CHECK_LAUNCH("center(k)");
    center_rows_kernel<<<1, 256, 0, s>>>(sp.to_v_w, sp.to_v_b, ps.Wv_c, ps.bv_c);
    SV_CHECK_LAUNCH("center(v)");
    transpose256_kernel<<<C, 256, 0, s>>>(ps.Wk_c, ps.Wk_cT);
    SV_CHECK_LAUNCH("transpose");
    if (sp.tq_to_q_w) {
      const float* ws[3] = {sp.tq_to_q_w, sp.tq_to_k_w, sp.tq_to_v_w};
      const float* bs[3] = {sp.tq_to_q_b, sp.tq_to_k_b, sp.tq_to_v_b};
      const float* lw[3] = {sp.tq_nq_w, sp.tq_nk_w, sp.tq_nv_w};
      const float* lb[3] = {sp.tq_nq_b, sp.tq_nk_b, sp.tq_nv_b};
      for (int j = 0; j < 3; ++j) {
        SV_TRY(dcopy(ws[j], ps.tq_qkv_w + (long)j * C * C, (long)C * C, s));
        SV_TRY(dcopy(bs[j], ps.tq_qkv_b + j * C, C, s));
        SV_TRY(dcopy(lw[j], ps.tq_ln_w + j * C, C, s));
        SV_TRY(dcopy(lb[j], ps.tq_ln_b + j * C, C, s));
      }
    }
    SV_TRY(dcopy(sp.cls0_w, ps.tw_w, (long)C * C, s));
    SV_TRY(dcopy(sp.reg0_w, ps.tw_w + (long)C * C, (long)C * C, s));
    SV_TRY(dcopy(sp.cls0_nw, ps.tw_ln_w, C, s)); SV_TRY(dcopy(sp.reg0_nw, ps.tw_ln_w + C, C, s));
    SV_TRY(dcopy(sp.cls0_nb, ps.tw_ln_b, C, s)); SV_TRY(dcopy(sp.reg0_nb, ps.tw_ln_b + C, C, s));
    SV_TRY(tc_prepare_stage(sp, ps.Wk_c, ps.bk_c, ps.Wv_c, ps.bv_c, ps.tc, s));
    if (slot::slot_tc_supported(d)) {
      const int F = d->dim_feedforward;
      SV_TRY(slot::slot_planes(sp.out_proj_w, C, C, ps.stc.out_proj, s));
      SV_TRY(slot::slot_planes(sp.to_q_w, C, C, ps.stc.to_q, s));
      SV_TRY(slot::slot_planes(ps.Wk_cT, C, C, ps.stc.wkT, s));
      SV_TRY(slot::slot_planes(ps.Wv_c, C, C, ps.stc.wv, s));
      SV_TRY(slot::slot_planes(sp.lin1_w, F, C, ps.stc.lin1, s));
      SV_TRY(slot::slot_planes(sp.lin2_w, C, F, ps.stc.lin2, s));
      SV_TRY(slot::slot_planes(ps.tw_w, 2 * C, C, ps.stc.tw, s));
      SV_TRY(slot::slot_planes(sp.cls1_w, C, C, ps.stc.cls1, s));
      SV_TRY(slot::slot_planes(sp.reg1_w, C, C, ps.stc.reg1, s));
      SV_TRY(slot::slot_planes(sp.logit_w, d->num_classes, C, ps.stc.logit, s));
      if (sp.tq_to_q_w) {
        const int TF = d->temporal_dim_feedforward;
        SV_TRY(slot::slot_planes(ps.tq_qkv_w, 3 * C, C, ps.stc.tqkv, s));
        SV_TRY(slot::slot_planes(sp.tq_lin1_w, TF, C, ps.stc.tlin1, s));
        SV_TRY(slot::slot_planes(sp.tq_lin2_w, C, TF, ps.stc.tlin2, s));
      }
    }
  }
  return SLOTVPS_OK;
}

int slotvps_head_forward(const slotvps_head_desc* d, const slotvps_stage_params* stages, const void* prepared,
                         const float* const* feats, const float* const* pos, const float* const* init_query,
                         float* cls_out, float* emb_out, float* const* fused_out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return slotvps_head_forward_ex(d, stages, prepared, feats, pos, init_query, cls_out, emb_out, fused_out, workspace, workspace_bytes,
                                 nullptr, stream);
}

int slotvps_head_forward_ex(const slotvps_head_desc* d, const slotvps_stage_params* stages, const void* prepared,
                            const float* const* feats, const float* const* pos, const float* const* init_query,
                            float* cls_out, float* emb_out, float* const* fused_out, void* workspace,
                            size_t workspace_bytes, const slotvps_head_opts* opts, void* stream) {
  SV_TRY(validate(d));
  slotvps_head_opts o;
  memset(&o, 0, sizeof(o));
  if (opts) o = *opts;
  SV_REQUIRE((o.feat_bn_scale == nullptr) == (o.feat_bn_shift == nullptr) && (o.feat_bn_scale == nullptr) == (o.rnorm_ss == nullptr),
             "feat_bn_scale, feat_bn_shift and rnorm_ss go together");
  SV_REQUIRE(stages && prepared && feats && init_query && cls_out && emb_out && fused_out && workspace, "null argument");
  SV_REQUIRE(d->pos_mode != 1 || pos != nullptr, "pos_mode 1 needs pos tensors");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  const int T = d->n_frames, N = d->n_slots, L = d->n_levels, S = n_stages_of(d);
  HeadWs w;
  if (head_ws_layout(d, workspace, workspace_bytes, &w) > workspace_bytes)
    return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  Prepared pr;
  prepared_layout(d, const_cast<void*>(prepared), &pr);
  // frame strides of the caller's fused_out / pos tensors must be uniform per level
  long fstride[SLOTVPS_MAX_LEVELS], pstride[SLOTVPS_MAX_LEVELS];
  for (int l = 0; l < L; ++l) {
    const bool skip = (o.skip_fused_mask >> l) & 1;
    fstride[l] = (T > 1 && !skip) ? (long)(fused_out[1 * L + l] - fused_out[l]) : 0;
    pstride[l] = (d->pos_mode == 1 && T > 1) ? (long)(pos[1 * L + l] - pos[l]) : 0;
    for (int t = 0; t < T; ++t) {
      SV_REQUIRE(skip || (fused_out[t * L + l] != nullptr && fused_out[t * L + l] == fused_out[l] + t * fstride[l]),
                 "fused_out frames of a level must be equally strided");
      if (d->pos_mode == 1) SV_REQUIRE(pos[t * L + l] == pos[l] + t * pstride[l], "pos frames of a level must be equally strided");
    }
  }
  if (const char* pz = getenv("SLOTVPS_POISON")) {          // debugging aid: pre-fill scratch buffers with a finite pattern
    const int m = atoi(pz);
    const int pv_ = getenv("SLOTVPS_POISON_BYTE") ? atoi(getenv("SLOTVPS_POISON_BYTE")) : 0x3C;
    int Pmax = 0;
    for (int l = 0; l < L; ++l) Pmax = max(Pmax, d->h[l] * d->w[l]);
    const size_t prow = (size_t)T * Pmax;
    if ((m & 1) && w.tc.planes) { cudaMemsetAsync(w.tc.planes, pv_, 4 * prow * C * 2, s); cudaMemsetAsync(w.tc.planes_alt, pv_, 4 * prow * C * 2, s); }
    if ((m & 2) && w.pg[0]) for (int i = 0; i < 2; ++i) cudaMemsetAsync(w.pg[i], 0x3C, (size_t)T * (w.hmax + w.wmax) * 112 * 4, s);
    if ((m & 4) && w.tk[0]) for (int i = 0; i < SLOTVPS_MAX_STAGES; ++i) cudaMemsetAsync(w.tk[i], 0x3C, (size_t)(w.hmax + w.wmax) * C * 4, s);
    if ((m & 8) && w.ftc.y) { cudaMemsetAsync(w.ftc.y, 0x3C, (size_t)T * (Pmax / 4 + 1) * C * 4, s); cudaMemsetAsync(w.ftc.in_planes, 0x3C, 2 * prow * CIN * 2, s); }
    if (m & 16) { cudaMemsetAsync(w.rs_k, 0x3C, prow * 4, s); cudaMemsetAsync(w.rs_v, 0x3C, prow * 4, s); cudaMemsetAsync(w.rs_k2, 0x3C, prow * 4, s); cudaMemsetAsync(w.rs_v2, 0x3C, prow * 4, s); }
    if ((m & 32) && w.tc.gplanes) cudaMemsetAsync(w.tc.gplanes, 0x3C, (size_t)T * 2 * 112 * C * 2, s);
    if (m & 64) { cudaMemsetAsync(w.Zpart, 0x3C, (size_t)148 * T * N * C * 4, s); cudaMemsetAsync(w.L, 0x3C, (size_t)T * N * T * N * 4, s); cudaMemsetAsync(w.tqkv, 0x3C, (size_t)T * N * 3 * C * 4, s); }
  }
  for (int t = 0; t < T; ++t) SV_TRY(dcopy(init_query[t], w.slots + (long)t * N * C, (long)N * C, s));
  const long cls_fs = (long)S * N * d->num_classes, emb_fs = (long)S * N * C;
  // Overlapped (two-stream) schedule when every level runs the tensor-core kernels end to end.
  bool overlap = d->kernel_path == 0;
  for (int l = 0; l < L; ++l) overlap = overlap && tc_supported(d, l) && d->heads_per_level[l] > 0;
  if (getenv("SLOTVPS_NO_OVERLAP") || g_prof_on) overlap = false;     // per-launch event timing needs one stream
  cudaStream_t sb = s;
  cudaEvent_t ev_attn[SLOTVPS_MAX_STAGES] = {nullptr};       // attention of stage i finished (main stream)
  int last_stage_of_level[SLOTVPS_MAX_LEVELS] = {0};
  Overlap& g_ov = overlap_next();
  if (overlap) {
    for (auto& o : g_ovs) SV_TRY(o.init());
    sb = g_ov.side;
    cudaEvent_t fork = g_ov.take();
    SV_CHECK_CUDA(cudaEventRecord(fork, s));
    SV_CHECK_CUDA(cudaStreamWaitEvent(sb, fork, 0));
  }
  int side_ctas = overlap ? SIDE_CTAS : 148;
  if (const char* e = overlap ? getenv("SLOTVPS_SIDE_CTAS") : nullptr) { const int v = atoi(e); if (v >= 16 && v <= 148) side_ctas = v; }
  int stage = 0;
  bool prev_planes = false;
  for (int l = 0; l < L; ++l) {
    const int h = d->h[l], wd = d->w[l], P = h * wd;
    const bool use_tc = d->kernel_path == 0 && tc_supported(d, l);
    const bool all_tc = use_tc;                               // no fp32 kernel touches pos at this level
    const bool fuse_tc = use_tc && (l == 0 || prev_planes);  // the coarse GEMM reads the previous level's planes
    const bool skip_out = (o.skip_fused_mask >> l) & 1;
    SV_REQUIRE(!skip_out || (fuse_tc && d->heads_per_level[l] > 0), "skip_fused_mask: level does not run the tensor-core path");
    SV_REQUIRE(o.rnorm_ss == nullptr || l != L - 1 || fuse_tc, "rnorm_ss needs the tensor-core level fusion on the finest level");
    // Separable-pos mode (x planes only, position terms from tables).  Measured on B200 at 1024x2048: level fusion
    // gets 0.15 ms/step faster (2 planes instead of 4) but the table loads make the statistics / attention epilogues
    // 0.19 ms/step slower, so it is opt-in (SLOTVPS_POS_SEP=1); pos == None always uses it (no tables needed).
    const bool pos_sep = use_tc && d->pos_mode != 1 && (d->pos_mode == 0 || (getenv("SLOTVPS_POS_SEP") != nullptr && N <= attn::NROW));
    TcWorkspace tcl = w.tc, tcp = w.tc;                      // this level's / the previous level's operand planes
    if (overlap) { tcl.planes = (l & 1) ? w.tc.planes_alt : w.tc.planes; tcp.planes = (l & 1) ? w.tc.planes : w.tc.planes_alt; }
    tcl.ytab = w.tc.ytab_l[l]; tcl.xtab = w.tc.xtab_l[l];      // per-level sine tables
    const float* pl = nullptr;
    long pls = 0;
    if (d->pos_mode == 1) { pl = pos[l]; pls = pstride[l]; }
    else if (d->pos_mode == 2 && !all_tc) {
      sine_pos_kernel<<<(unsigned)(((long)C * P + 255) / 256), 256, 0, s>>>(w.pos[l], h, wd);
      SV_CHECK_LAUNCH("sine_pos");
      pl = w.pos[l]; pls = 0;
    }
    if (fuse_tc) {
      // ---- level fusion on the tensor pipe (side stream when overlapped); the epilogue also emits the operand planes ----
      cudaStream_t s = sb;                                  // shadows the main stream inside this block
      if (overlap && l >= 2) SV_CHECK_CUDA(cudaStreamWaitEvent(sb, ev_attn[last_stage_of_level[l - 2]], 0));   // planes[l&1] are free again
      const long rows = (long)T * P;
      Ptr8 src;
      for (int t = 0; t < SLOTVPS_MAX_FRAMES; ++t) src.p[t] = t < T ? feats[t * L + l] : nullptr;
      bool vec4 = P % 4 == 0;
      for (int t = 0; t < T; ++t) vec4 = vec4 && ((uintptr_t)src.p[t] & 15) == 0;
      if (vec4) split_in4_kernel<<<dim3(ceil_div(P, 128), T), 256, 0, s>>>(src, w.ftc.in_planes, rows, P);
      else split_in_kernel<<<dim3(ceil_div(P, 32), T), 256, 0, s>>>(src, w.ftc.in_planes, rows, P);
      SV_CHECK_LAUNCH("split_in");
      fuse::Params prm;
      memset(&prm, 0, sizeof(prm));
      if (l > 0) {
        const int Pp = d->h[l - 1] * d->w[l - 1];
        const long rp = (long)T * Pp;
        prm.rows = (int)rp; prm.P = Pp; prm.w = d->w[l - 1]; prm.h = d->h[l - 1]; prm.ksub = 4; prm.a_lo_row = (int)rp;
        prm.y_out = w.ftc.y; prm.a_split = 1;
        SV_TRY(fuse_tc_launch(tcp.planes, 2 * rp, (int)rp, C, pr.ftc.wa, prm, s, side_ctas));
        memset(&prm, 0, sizeof(prm));
      }
      prm.rows = (int)rows; prm.P = P; prm.w = wd; prm.h = h; prm.ksub = 2; prm.a_lo_row = (int)rows;
      prm.bias = l > 0 ? pr.conv_b : pr.conv_b0; prm.y_in = l > 0 ? w.ftc.y : nullptr;
      prm.out = skip_out ? nullptr : fused_out[l]; prm.out_bs = fstride[l];
      if (o.rnorm_ss && l == L - 1) {
        prm.bn_sc = o.feat_bn_scale; prm.bn_sh = o.feat_bn_shift; prm.ss_out = o.rnorm_ss;
      }
      prm.planes = tcl.planes; prm.plane_stride = rows; prm.x_planes_only = pos_sep ? 1 : 0;
      { static const int ts = getenv("SLOTVPS_FUSE_TMA_STORE") ? atoi(getenv("SLOTVPS_FUSE_TMA_STORE")) : 1; prm.tma_store = ts; }
      if (d->pos_mode == 1) { prm.pos = pos[l]; prm.pos_bs = pstride[l]; }
      else if (d->pos_mode == 2) {
        pos_tab_kernel<<<ceil_div(128 * (h + wd), 256), 256, 0, s>>>(tcl.ytab, tcl.xtab, h, wd, w.tc.ytabT_l[l]);
        SV_CHECK_LAUNCH("pos_tab");
        prm.ytab = tcl.ytab; prm.xtab = tcl.xtab; prm.ytabT = w.tc.ytabT_l[l];
      }
      SV_TRY(fuse_tc_launch(w.ftc.in_planes, 2 * rows, (int)rows, CIN, l > 0 ? pr.ftc.wb : pr.ftc.w0, prm, s, side_ctas));
    } else {
      for (int t = 0; t < T; ++t)
        SV_TRY(level_fuse_frame(l > 0 ? fused_out[t * L + l - 1] : nullptr, feats[t * L + l], pr.conv_w, l > 0 ? pr.conv_b : pr.conv_b0, pr.W0,
                                fused_out[t * L + l], h, wd, w.ybuf + (long)t * C * (P / 4 + 1), s));
      if (use_tc && d->heads_per_level[l] > 0) {
        SV_TRY(tc_split_level(fused_out[l], fstride[l], pl, pls, d->pos_mode == 2 && all_tc, tcl, T, h, wd, s));
        if (pos_sep && d->pos_mode == 2 && !all_tc) {          // tables for the key statistics even when pos was materialised
          pos_tab_kernel<<<ceil_div(128 * (h + wd), 256), 256, 0, s>>>(tcl.ytab, tcl.xtab, h, wd);
          SV_CHECK_LAUNCH("pos_tab");
        }
      }
    }
    prev_planes = fuse_tc || (use_tc && d->heads_per_level[l] > 0);
    // ---- per-stage pixel inputs; in overlapped mode the statistics of every stage of the level are queued now ----
    StagePix px[SLOTVPS_MAX_STAGES];
    for (int j = 0; j < d->heads_per_level[l]; ++j) {
      const int st = stage + j;
      StagePix& q = px[j];
      q.tc = tcl;
      q.pg = w.pg[st & 1];
      if (pos_sep) {
        q.ps.enabled = 1; q.ps.w = wd; q.ps.h = h;
        if (d->pos_mode == 2) {                               // key tables of this stage (same stream as the planes' producer)
          cudaStream_t s = sb;
          float* tky = w.tk[st];
          float* tkx = w.tk[st] + (long)h * C;
          GemmArgs g;
          g.A = tcl.ytab; g.a_ms = 1; g.a_ks = h;
          g.B = pr.st[st].Wk_c; g.b_ks = 1; g.b_ns = C;
          g.Cm = tky; g.c_ms = C; g.c_ns = 1;
          g.M = h; g.N = C; g.K = 128;
          SV_TRY(sgemm(g, s));
          g.A = tcl.xtab; g.a_ks = wd;
          g.B = pr.st[st].Wk_c + 128;
          g.Cm = tkx; g.M = wd;
          SV_TRY(sgemm(g, s));
          q.ps.tky = tky; q.ps.tkx = tkx;
        }
      }
      q.rs_k = (overlap && (st & 1)) ? w.rs_k2 : w.rs_k;
      q.rs_v = (overlap && (st & 1)) ? w.rs_v2 : w.rs_v;
      if (overlap) {
        q.overlapped = true;
        q.ready = g_ov.take(); q.done = g_ov.take();
        ev_attn[st] = q.done;
        cudaStream_t s = sb;
        if (st >= 2) SV_CHECK_CUDA(cudaStreamWaitEvent(sb, ev_attn[st - 2], 0));     // rs buffers of parity st&1 are free again
        SV_TRY(tc_stats(pr.st[st].tc, tcl, pr.st[st].bk_c, pr.st[st].bv_c, q.rs_k, q.rs_v, T, P, s, side_ctas, q.ps));
        SV_CHECK_CUDA(cudaEventRecord(q.ready, sb));
      }
    }
    for (int j = 0; j < d->heads_per_level[l]; ++j, ++stage) {
      const bool temporal = (d->temporal_mask >> stage) & 1;
      SV_TRY(run_stage(d, stages[stage], pr.st[stage], w, px[j], fused_out[l], fstride[l], pl, pls, h, wd, temporal, use_tc,
                       cls_out + (long)stage * N * d->num_classes, cls_fs, emb_out + (long)stage * N * C, emb_fs,
                       o.stage_slots_in ? o.stage_slots_in + (long)stage * T : nullptr, s));
    }
    last_stage_of_level[l] = stage - 1;
  }
  g_last_head_alt = overlap && ((L - 1) & 1);
  if (overlap) {                                            // join: fused_out / planes written on the side stream
    cudaEvent_t join = g_ov.take();
    SV_CHECK_CUDA(cudaEventRecord(join, sb));
    SV_CHECK_CUDA(cudaStreamWaitEvent(s, join, 0));
  }
  return SLOTVPS_OK;
}

int slotvps_level_fuse(const float* prev, const float* x, const float* conv_w, const float* conv_b, float* out, int h, int w,
                       float* scratch, void* stream) {
  SV_REQUIRE(x && conv_w && conv_b && out && h > 0 && w > 0, "bad argument");
  SV_REQUIRE(scratch != nullptr, "scratch of 256*max(128, (h/2)*(w/2)) floats required");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  if (!prev) {
    fold_w0_kernel<<<ceil_div(C * CIN, 256), 256, 0, s>>>(conv_w, scratch);
    SV_CHECK_LAUNCH("fold_w0");
    return level_fuse_frame(nullptr, x, conv_w, conv_b, scratch, out, h, w, nullptr, s);
  }
  SV_REQUIRE(h % 2 == 0 && w % 2 == 0, "level must be 2x the previous");
  return level_fuse_frame(prev, x, conv_w, conv_b, nullptr, out, h, w, scratch, s);
}

// ---- tracker --------------------------------------------------------------------------------------------
namespace {
constexpr int TRACK_MAX_CAPACITY = 4000;            // 12 B of shared memory per bank row in track_assign_kernel
struct TrackLayout { track::State* st; float* bank; float* y_cur; float* y_bank; float* lik; int* mid; size_t bytes; };
TrackLayout track_layout(void* state, int capacity, int n_slots) {
  Arena a(state, (size_t)-1);
  TrackLayout t;
  t.st = (track::State*)a.take<char>(256);
  t.bank = a.take<float>((size_t)capacity * C);
  t.y_cur = a.take<float>((size_t)n_slots * C);
  t.y_bank = a.take<float>((size_t)capacity * C);
  t.lik = a.take<float>(n_slots);
  t.mid = a.take<int>(n_slots);
  t.bytes = align_up(a.off);
  return t;
}
}  // namespace

int slotvps_track_scores(const float* fc_w, const float* fc_b, int num_fcs, const float* x_query, int k,
                         const float* ref_x_query, int m, float* match_score, void* workspace, size_t workspace_bytes,
                         void* stream) {
  SV_REQUIRE(x_query && ref_x_query && match_score && workspace, "null argument");
  SV_REQUIRE(k > 0 && m > 0 && m <= TRACK_MAX_CAPACITY && num_fcs >= 0 && (num_fcs == 0 || (fc_w && fc_b)), "bad argument");
  if (workspace_bytes < (size_t)(k + m) * C * sizeof(float)) return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  float* ya = (float*)workspace; float* yb = ya + (size_t)k * C;
  track::track_fc_kernel<<<ceil_div(k, track::ROWS), 256, 0, s>>>(x_query, nullptr, nullptr, k, fc_w, fc_b, num_fcs, ya);
  SV_CHECK_LAUNCH("track_fc");
  track::track_fc_kernel<<<ceil_div(m, track::ROWS), 256, 0, s>>>(ref_x_query, nullptr, nullptr, m, fc_w, fc_b, num_fcs, yb);
  SV_CHECK_LAUNCH("track_fc");
  track::track_score_kernel<<<k, 256, (size_t)(1 + m) * sizeof(float), s>>>(ya, yb, nullptr, k, nullptr, m, nullptr, nullptr, match_score, 1 + m);
  SV_CHECK_LAUNCH("track_score");
  return SLOTVPS_OK;
}

int slotvps_track_state_bytes(int capacity, int n_slots, size_t* bytes) {
  SV_REQUIRE(bytes && capacity > 0 && capacity <= TRACK_MAX_CAPACITY && n_slots > 0 && n_slots <= 1024, "bad argument");
  *bytes = track_layout(nullptr, capacity, n_slots).bytes;
  return SLOTVPS_OK;
}

int slotvps_track_reset(void* state, size_t state_bytes, void* stream) {
  SV_REQUIRE(state && state_bytes >= 256, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  track::track_reset_kernel<<<1, 1, 0, s>>>((track::State*)state);
  SV_CHECK_LAUNCH("track_reset");
  return SLOTVPS_OK;
}

int slotvps_track_step(const float* fc_w, const float* fc_b, int num_fcs, const float* embedding, const int32_t* fusion_meta,
                       int n_slots, void* state, size_t state_bytes, int capacity, int32_t* track_out, void* stream) {
  SV_REQUIRE(embedding && fusion_meta && state && track_out, "null argument");
  SV_REQUIRE(capacity > 0 && capacity <= TRACK_MAX_CAPACITY && n_slots > 0 && n_slots <= 1024, "bad argument");
  SV_REQUIRE(num_fcs >= 0 && (num_fcs == 0 || (fc_w && fc_b)), "bad argument");
  const TrackLayout t = track_layout(state, capacity, n_slots);
  if (state_bytes < t.bytes) return fail(SLOTVPS_EWORKSPACE, "tracker state too small%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  // Both FC launches and the score launch size their grids for the maxima and exit on the device-side counts
  // (K' = meta[0], bank rows = State::count), so no count is read back; on the first frame of a video the bank
  // is empty and the assign kernel ignores lik / mid.
  track::track_fc_kernel<<<ceil_div(n_slots, track::ROWS), 256, 0, s>>>(embedding, fusion_meta + 4, fusion_meta, n_slots, fc_w, fc_b, num_fcs, t.y_cur);
  SV_CHECK_LAUNCH("track_fc");
  track::track_fc_kernel<<<ceil_div(capacity, track::ROWS), 256, 0, s>>>(t.bank, nullptr, &t.st->count, capacity, fc_w, fc_b, num_fcs, t.y_bank);
  SV_CHECK_LAUNCH("track_fc");
  track::track_score_kernel<<<n_slots, 256, (size_t)(1 + capacity) * sizeof(float), s>>>(t.y_cur, t.y_bank, fusion_meta, n_slots, &t.st->count, capacity,
                                                                                        t.lik, t.mid, nullptr, 0);
  SV_CHECK_LAUNCH("track_score");
  if ((size_t)capacity * 12 + 4200 > 48 * 1024) SV_TRY(ensure_dyn_smem((const void*)track::track_assign_kernel, (size_t)capacity * 12));
  track::track_assign_kernel<<<1, 256, (size_t)capacity * 12, s>>>(t.st, t.bank, capacity, embedding, fusion_meta, n_slots, t.lik, t.mid, track_out);
  SV_CHECK_LAUNCH("track_assign");
  return SLOTVPS_OK;
}

// ---- consumers of the id map ------------------------------------------------------------------------------
int slotvps_semantic_argmax(const float* fcn_output, int n_classes, int h, int w, int H, int W, int64_t* out, void* stream) {
  SV_REQUIRE(fcn_output && out, "null argument");
  SV_REQUIRE(n_classes > 0 && n_classes <= unify::MAXSEM && h > 0 && w > 0 && H > 0 && W > 0, "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  const long HW = (long)H * W;
  const int grid = (int)((HW + 255) / 256 < 148 * 8 ? (HW + 255) / 256 : 148 * 8);
  unify::semantic_argmax_kernel<<<grid, 256, 0, s>>>(fcn_output, n_classes, h, w, H, W, (long long*)out);
  SV_CHECK_LAUNCH("semantic_argmax");
  return SLOTVPS_OK;
}

namespace {
struct UnifyLayout { unify::State* st; unsigned int* hist; unsigned char* luts; size_t bytes; };
UnifyLayout unify_layout(void* ws) {
  Arena a(ws, (size_t)-1);
  UnifyLayout u;
  u.st = (unify::State*)a.take<char>(256);
  u.hist = a.take<unsigned int>((size_t)unify::MAXID * unify::MAXSEM);
  u.luts = a.take<unsigned char>(3 * unify::MAXID);
  u.bytes = align_up(a.off);
  return u;
}
}  // namespace

int slotvps_unify_workspace_bytes(size_t* bytes) {
  SV_REQUIRE(bytes != nullptr, "null out pointer");
  *bytes = unify_layout(nullptr).bytes;
  return SLOTVPS_OK;
}

int slotvps_unify_reset(void* workspace, size_t workspace_bytes, void* stream) {
  SV_REQUIRE(workspace && workspace_bytes >= unify_layout(nullptr).bytes, "bad workspace");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  unify::unify_reset_kernel<<<1, 1, 0, s>>>(unify_layout(workspace).st);
  SV_CHECK_LAUNCH("unify_reset");
  return SLOTVPS_OK;
}

int slotvps_unify_pan_result(const int64_t* seg, const int64_t* pan, const int32_t* cls_inds, int n_inst, const int32_t* obj_ids,
                             int n_obj, int H, int W, int id_last_stuff, int stuff_area_limit, uint8_t* pan_2ch, int32_t* status,
                             void* workspace, size_t workspace_bytes, void* stream) {
  SV_REQUIRE(seg && pan && pan_2ch && workspace, "null argument");
  SV_REQUIRE(H > 0 && W > 0 && n_inst >= 0 && n_inst <= unify::MAXID && n_obj >= 0 && n_obj <= unify::MAXID, "bad shape");
  SV_REQUIRE((n_inst == 0 || cls_inds) && (n_obj == 0 || obj_ids), "null id arrays");
  SV_REQUIRE(id_last_stuff >= 0 && id_last_stuff < unify::MAXID - 1 && stuff_area_limit >= 0, "bad argument");
  const UnifyLayout u = unify_layout(workspace);
  if (workspace_bytes < u.bytes) return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  const long HW = (long)H * W;
  SV_CHECK_CUDA(cudaMemsetAsync(u.hist, 0, sizeof(unsigned int) * unify::MAXID * unify::MAXSEM, s));
  const long chunks = (HW + 2047) / 2048;
  unify::unify_hist_kernel<<<(int)(chunks < 148 * 4 ? chunks : 148 * 4), 256, 0, s>>>((const long long*)seg, (const long long*)pan, HW, u.hist, u.st);
  SV_CHECK_LAUNCH("unify_hist");
  unify::unify_decide_kernel<<<1, unify::MAXID, 0, s>>>(u.hist, cls_inds, n_inst, n_obj > 0 ? obj_ids : nullptr, n_obj, id_last_stuff,
                                                       (unsigned int)stuff_area_limit, u.st, u.luts);
  SV_CHECK_LAUNCH("unify_decide");
  const long quads = (HW + 3) / 4;
  unify::unify_write_kernel<<<(int)((quads + 255) / 256 < 148 * 8 ? (quads + 255) / 256 : 148 * 8), 256, 0, s>>>((const long long*)pan, HW, u.luts, pan_2ch);
  SV_CHECK_LAUNCH("unify_write");
  if (status) SV_CHECK_CUDA(cudaMemcpyAsync(status, u.st, 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  return SLOTVPS_OK;
}

int slotvps_sine_pos(float* out, int h, int w, void* stream) {
  SV_REQUIRE(out && h > 0 && w > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  sine_pos_kernel<<<(unsigned)(((long)C * h * w + 255) / 256), 256, 0, s>>>(out, h, w);
  SV_CHECK_LAUNCH("sine_pos");
  return SLOTVPS_OK;
}

// ---- mask logits ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_prep_kernel(const float* __restrict__ emb, const float* __restrict__ bw,
                                                        const float* __restrict__ bb, const float* __restrict__ bm,
                                                        const float* __restrict__ bv, const float* __restrict__ fg,
                                                        float* __restrict__ sc, float* __restrict__ sh, float* __restrict__ e2,
                                                        float* __restrict__ dn, float* __restrict__ aff, int N) {
  // block 0..N-1: one slot row each (warp 0 does the row dot); also (block 0) the folded BN vectors
  const int n = blockIdx.x, c = threadIdx.x;
  const float s = bw[c] / sqrtf(bv[c] + BN_EPS);
  const float t = bb[c] - bm[c] * s;
  if (n == 0) {
    sc[c] = s; sh[c] = t;
    if (c == 0) { float sg = fg[0] / sqrtf(fg[3] + BN_EPS); aff[0] = sg; aff[1] = fg[1] - fg[2] * sg; }
  }
  const float e = emb[(long)n * C + c];
  e2[(long)n * C + c] = e * s;
  __shared__ float red[8];
  float v = warp_sum(e * t);
  if ((c & 31) == 0) red[c >> 5] = v;
  __syncthreads();
  if (c == 0) { float a = 0.f; for (int i = 0; i < 8; ++i) a += red[i]; dn[n] = a; }
}

// eval-mode BatchNorm as a per-channel affine: scale = w / sqrt(var + eps), shift = b - mean * scale (same expressions as mask_prep_kernel)
__global__ void __launch_bounds__(256) bn_fold_kernel(const float* __restrict__ bw, const float* __restrict__ bb, const float* __restrict__ bm,
                                                      const float* __restrict__ bv, float* __restrict__ sc, float* __restrict__ sh, int n) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n) return;
  const float s = bw[c] / sqrtf(bv[c] + BN_EPS);
  sc[c] = s; sh[c] = bb[c] - bm[c] * s;
}
int slotvps_fold_batchnorm(const float* w, const float* b, const float* mean, const float* var, int n, float* scale, float* shift, void* stream) {
  SV_REQUIRE(w && b && mean && var && scale && shift && n > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  bn_fold_kernel<<<ceil_div(n, 256), 256, 0, s>>>(w, b, mean, var, scale, shift, n);
  SV_CHECK_LAUNCH("bn_fold");
  return SLOTVPS_OK;
}

int slotvps_mask_logits_workspace_bytes(int n_slots, int h, int w, size_t* bytes) {
  SV_REQUIRE(bytes && n_slots > 0 && h > 0 && w > 0, "bad argument");
  Arena a(nullptr, (size_t)-1);
  a.take<float>(C); a.take<float>(C); a.take<float>((size_t)n_slots * C); a.take<float>(n_slots); a.take<float>(4);
  a.take<float>((size_t)h * w);
  a.take<__half>((size_t)2 * 112 * C);
  *bytes = align_up(a.off);
  return SLOTVPS_OK;
}

int slotvps_mask_logits(const float* feat, const float* emb, const float* bw, const float* bb, const float* bm, const float* bv,
                        const float* fg_bn, float* out, int n_slots, int h, int w, void* workspace, size_t workspace_bytes,
                        void* stream) {
  SV_REQUIRE(feat && emb && bw && bb && bm && bv && fg_bn && out && workspace, "null argument");
  SV_REQUIRE(n_slots > 0 && h > 0 && w > 0, "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  const int P = h * w;
  Arena a(workspace, workspace_bytes);
  float* sc = a.take<float>(C); float* sh = a.take<float>(C);
  float* e2 = a.take<float>((size_t)n_slots * C); float* dn = a.take<float>(n_slots); float* aff = a.take<float>(4);
  float* rn = a.take<float>(P);
  if (!a.ok()) return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  mask_prep_kernel<<<n_slots, 256, 0, s>>>(emb, bw, bb, bm, bv, fg_bn, sc, sh, e2, dn, aff, n_slots);
  SV_CHECK_LAUNCH("mask_prep");
  if (P % 4 == 0 && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)rn & 15) == 0) feat_rnorm4_kernel<<<ceil_div(P, 256), 256, 0, s>>>(feat, sc, sh, rn, P);
  else feat_rnorm_kernel<<<ceil_div(P, 256), 256, 0, s>>>(feat, sc, sh, rn, P);
  SV_CHECK_LAUNCH("feat_rnorm");
  GemmArgs g;
  g.A = e2; g.a_ms = C; g.a_ks = 1;
  g.B = feat; g.b_ks = P; g.b_ns = 1;
  g.Cm = out; g.c_ms = P; g.c_ns = 1;
  g.M = n_slots; g.N = P; g.K = C;
  g.bias = dn; g.bias_mode = 1;
  g.col_scale = rn; g.affine = aff;
  return sgemm(g, s);
}

// Mask logits straight from the operand planes a preceding slotvps_head_forward left in `head_workspace`
// (finest level, frame `frame`).  Returns SLOTVPS_EUNSUPPORTED when that head call did not take the
// tensor-core path for the finest level (caller then uses slotvps_mask_logits on the fp32 feature).
int slotvps_head_mask_logits(const slotvps_head_desc* d, void* head_workspace, size_t head_workspace_bytes, int frame, const float* feat,
                             const float* emb, const float* bw, const float* bb, const float* bm, const float* bv, const float* fg_bn,
                             float* out, void* workspace, size_t workspace_bytes, void* stream) {
  return slotvps_head_mask_logits_ex(d, head_workspace, head_workspace_bytes, frame, feat, nullptr, emb, bw, bb, bm, bv, fg_bn, out, workspace,
                                     workspace_bytes, stream);
}

// feat: the fp32 feature (per-pixel norm computed here), or NULL with rnorm_ss = the per-pixel squared norms the level-fusion
// epilogue of the preceding slotvps_head_forward_ex accumulated ([T][P]; frame `frame` is used).
int slotvps_head_mask_logits_ex(const slotvps_head_desc* d, void* head_workspace, size_t head_workspace_bytes, int frame, const float* feat,
                                const float* rnorm_ss, const float* emb, const float* bw, const float* bb, const float* bm, const float* bv,
                                const float* fg_bn, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  SV_TRY(validate(d));
  SV_REQUIRE(head_workspace && emb && bw && bb && bm && bv && fg_bn && out && workspace, "null argument");
  SV_REQUIRE((feat != nullptr) != (rnorm_ss != nullptr), "exactly one of feat / rnorm_ss");
  SV_REQUIRE(frame >= 0 && frame < d->n_frames, "frame out of range");
  const int l = d->n_levels - 1, N = d->n_slots, h = d->h[l], w = d->w[l], P = h * w;
  if (!(d->kernel_path == 0 && tc_supported(d, l))) return fail(SLOTVPS_EUNSUPPORTED, "planes unavailable%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  HeadWs hw;
  if (head_ws_layout(d, head_workspace, head_workspace_bytes, &hw) > head_workspace_bytes) return fail(SLOTVPS_EWORKSPACE, "head workspace too small%s%s");
  Arena a(workspace, workspace_bytes);
  float* sc = a.take<float>(C); float* sh = a.take<float>(C);
  float* e2 = a.take<float>((size_t)N * C); float* dn = a.take<float>(N); float* aff = a.take<float>(4);
  float* rn = a.take<float>(P);
  __half* ep = a.take<__half>((size_t)2 * mask::NROW * C);
  if (!a.ok()) return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  mask_prep_kernel<<<N, 256, 0, s>>>(emb, bw, bb, bm, bv, fg_bn, sc, sh, e2, dn, aff, N);
  SV_CHECK_LAUNCH("mask_prep");
  const float* rnp = rn;
  if (feat) {
    if (P % 4 == 0 && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)rn & 15) == 0) feat_rnorm4_kernel<<<ceil_div(P, 256), 256, 0, s>>>(feat, sc, sh, rn, P);
    else feat_rnorm_kernel<<<ceil_div(P, 256), 256, 0, s>>>(feat, sc, sh, rn, P);
    SV_CHECK_LAUNCH("feat_rnorm");
  } else {
    rnp = rnorm_ss + (long)frame * P;                        // four partial sums, (long)d->n_frames * P apart
  }
  const long rows = (long)d->n_frames * P;
  // the finest level used the alternate plane set iff the head call ran the overlapped schedule (recorded by it)
  const __half* planes = g_last_head_alt ? hw.tc.planes_alt : hw.tc.planes;
  // slots are independent here, so N > 104 simply runs the kernel once per group of <= 104 slots
  const int groups = ceil_div(N, mask::NROW), base = N / groups, extra = N % groups;
  for (int g = 0, n0 = 0; g < groups; ++g) {
    const int ng = base + (g < extra ? 1 : 0);
    g_planes_kernel<<<(unsigned)(((long)mask::NROW * C + 255) / 256), 256, 0, s>>>(e2, ep, ng, 1, n0, N);
    SV_CHECK_LAUNCH("g_planes");
    SV_TRY(mask_tc_launch(planes, 2 * rows, rows, (long)frame * P, ep, dn + n0, rnp, aff, out + (long)n0 * P, ng, P, s, feat ? 0 : (int)rows));
    n0 += ng;
  }
  return SLOTVPS_OK;
}

// ---- panoptic fusion ----------------------------------------------------------------------------------
struct FuseWs {
  FuseState* st;
  unsigned int* pair;
  unsigned int* cand;        // per pixel: the (at most two) thing candidates with prob >= pixel_threshold
  unsigned short* ids;
};
static size_t fuse_ws_layout(int N, int H, int W, void* base, size_t cap, FuseWs* out) {
  Arena a(base, cap);
  FuseWs w;
  w.st = a.take<FuseState>(1);
  w.pair = a.take<unsigned int>((size_t)N * N);
  w.cand = a.take<unsigned int>((size_t)H * W);
  w.ids = a.take<unsigned short>((size_t)H * W);
  if (out) *out = w;
  return align_up(a.off);
}
int slotvps_fusion_workspace_bytes(int n_slots, int H, int W, size_t* bytes) {
  SV_REQUIRE(bytes && n_slots > 0 && n_slots <= FUSE_MAXN && H > 0 && W > 0, "bad argument");
  *bytes = fuse_ws_layout(n_slots, H, W, nullptr, (size_t)-1, nullptr);
  return SLOTVPS_OK;
}
static int fuse_iterate(const slotvps_fusion_cfg* cfg, const float* pred_masks, int N, int h, int w, int H, int W, int64_t* panoptic,
                        int32_t* meta, float* masks_out, int masks_cap, const FuseWs& ws, int iters, cudaStream_t s) {
  const long HW = (long)H * W;
  const int grid = (int)((HW + 255) / 256 < 148 * 16 ? (HW + 255) / 256 : 148 * 16);
  const bool x4 = (H == 4 * h && W == 4 * w);              // the shipped case: 1/4-resolution masks
  const long nblk = (long)h * w;
  const int grid4 = (int)((nblk + 255) / 256 < 148 * 8 ? (nblk + 255) / 256 : 148 * 8);
  FilterArgs fa{cfg->stuff_num, (unsigned)cfg->small_area, N, meta};
  for (int it = 0; it < iters; ++it) {                      // every pass exits at once when the fixed point was reached
    if (x4) fuse_argmax4_kernel<<<grid4, 256, 0, s>>>(pred_masks, h, w, ws.st, ws.cand, ws.ids, fa);
    else fuse_argmax_kernel<<<grid, 256, 0, s>>>(pred_masks, h, w, H, W, ws.st, ws.cand, ws.ids, fa);
    SV_CHECK_LAUNCH("fuse_argmax");
  }
  fuse_relabel_kernel<<<grid, 256, 0, s>>>(ws.st, ws.ids, HW, (long long*)panoptic);
  SV_CHECK_LAUNCH("fuse_relabel");
  if (masks_out && masks_cap > 0) {
    fuse_masks_kernel<<<grid, 256, 0, s>>>(pred_masks, h, w, H, W, ws.st, ws.cand, masks_out, masks_cap);
    SV_CHECK_LAUNCH("fuse_masks");
  }
  return SLOTVPS_OK;
}

int slotvps_panoptic_fuse(const slotvps_fusion_cfg* cfg, const float* pred_logits, const float* pred_masks, int N, int h, int w,
                          int H, int W, int64_t* panoptic, int32_t* meta, float* masks_out, int masks_cap, void* workspace,
                          size_t workspace_bytes, void* stream) {
  SV_REQUIRE(cfg && pred_logits && pred_masks && panoptic && meta && workspace, "null argument");
  SV_REQUIRE(N > 0 && N <= FUSE_MAXN && h > 0 && w > 0 && H > 0 && W > 0, "bad shape");
  SV_REQUIRE(cfg->num_classes >= 2 && cfg->stuff_num >= 0 && cfg->max_iters >= 1, "bad config");
  // the exact two-pass mask_removal keeps at most two candidates per pixel: needs pixel_threshold > 1/3
  SV_REQUIRE(cfg->pixel_threshold > 1.f / 3.f, "pixel_threshold must be > 1/3");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  FuseWs ws;
  if (fuse_ws_layout(N, H, W, workspace, workspace_bytes, &ws) > workspace_bytes)
    return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  const long HW = (long)H * W;
  const int grid = (int)((HW + 255) / 256 < 148 * 16 ? (HW + 255) / 256 : 148 * 16);
  // logits_width == num_classes: the last column is "no object" (:688-691); num_classes - 1 columns: no such test (:692-693)
  const int width = cfg->logits_width > 0 ? cfg->logits_width : cfg->num_classes;
  SV_REQUIRE(width == cfg->num_classes || width == cfg->num_classes - 1, "logits_width must be num_classes or num_classes - 1");
  fuse_select_kernel<<<1, FUSE_MAXN, 0, s>>>(pred_logits, N, width, cfg->stuff_num, cfg->threshold, width == cfg->num_classes ? 1 : 0, ws.st,
                                             ws.pair, meta);
  SV_CHECK_LAUNCH("fuse_select");
  const bool x4 = (H == 4 * h && W == 4 * w);
  const long nblk = (long)h * w;
  const int grid4 = (int)((nblk + 255) / 256 < 148 * 8 ? (nblk + 255) / 256 : 148 * 8);
  if (x4) fuse_count4_kernel<<<grid4, 256, 0, s>>>(pred_masks, h, w, cfg->pixel_threshold, cfg->fraction_threshold, ws.st, ws.pair, ws.cand);
  else fuse_count_kernel<<<grid, 256, 0, s>>>(pred_masks, h, w, H, W, cfg->pixel_threshold, cfg->fraction_threshold, ws.st, ws.pair, ws.cand);
  SV_CHECK_LAUNCH("fuse_count");
  return fuse_iterate(cfg, pred_masks, N, h, w, H, W, panoptic, meta, masks_out, masks_cap, ws, cfg->max_iters, s);
}

// Continue the small-segment filter loop of a slotvps_panoptic_fuse call whose meta[3] came back 0 (not converged within
// cfg->max_iters passes): `iters` more argmax / filter passes from the device state in `workspace`, then relabel.
int slotvps_panoptic_fuse_resume(const slotvps_fusion_cfg* cfg, const float* pred_masks, int N, int h, int w, int H, int W,
                                 int64_t* panoptic, int32_t* meta, float* masks_out, int masks_cap, void* workspace,
                                 size_t workspace_bytes, int iters, void* stream) {
  SV_REQUIRE(cfg && pred_masks && panoptic && meta && workspace && iters >= 1, "bad argument");
  SV_REQUIRE(N > 0 && N <= FUSE_MAXN && h > 0 && w > 0 && H > 0 && W > 0, "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  FuseWs ws;
  if (fuse_ws_layout(N, H, W, workspace, workspace_bytes, &ws) > workspace_bytes)
    return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  return fuse_iterate(cfg, pred_masks, N, h, w, H, W, panoptic, meta, masks_out, masks_cap, ws, iters, s);
}

// ---- Panoptic Retriever attention alone (per-kernel parity entry point) ----------------------------------
int slotvps_slot_attention_workspace_bytes(int n_slots, int h, int w, size_t* bytes) {
  SV_REQUIRE(bytes && n_slots > 0 && n_slots <= 512 && h > 0 && w > 0, "bad argument");
  slotvps_head_desc d;
  memset(&d, 0, sizeof(d));
  d.n_frames = 1; d.n_slots = n_slots; d.n_levels = 1; d.heads_per_level[0] = 1; d.h[0] = h; d.w[0] = w;
  d.num_classes = 20; d.dim_feedforward = 2048; d.temporal_dim_feedforward = 1024; d.nhead = 8;
  size_t a = head_ws_layout(&d, nullptr, (size_t)-1, nullptr), b = prepared_layout(&d, nullptr, nullptr);
  *bytes = a + b + 4096;
  return SLOTVPS_OK;
}
int slotvps_slot_attention(const slotvps_stage_params* sp, const float* slots_p, const float* x, const float* pos, float* out,
                           int N, int h, int wd, int kernel_path, void* workspace, size_t workspace_bytes, void* stream) {
  SV_REQUIRE(sp && slots_p && x && out && workspace, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  slotvps_head_desc d;
  memset(&d, 0, sizeof(d));
  d.n_frames = 1; d.n_slots = N; d.n_levels = 1; d.heads_per_level[0] = 1; d.h[0] = h; d.w[0] = wd;
  d.num_classes = 20; d.dim_feedforward = 2048; d.temporal_dim_feedforward = 1024; d.nhead = 8;
  d.kernel_path = kernel_path;
  SV_TRY(validate(&d));
  const size_t pb = prepared_layout(&d, nullptr, nullptr);
  if (pb + head_ws_layout(&d, nullptr, (size_t)-1, nullptr) > workspace_bytes) return fail(SLOTVPS_EWORKSPACE, "workspace too small%s%s");
  Prepared pr;
  prepared_layout(&d, workspace, &pr);
  HeadWs w;
  head_ws_layout(&d, (char*)workspace + pb, workspace_bytes - pb, &w);
  PreparedStage& ps = pr.st[0];
  center_rows_kernel<<<1, 256, 0, s>>>(sp->to_k_w, sp->to_k_b, ps.Wk_c, ps.bk_c);
  SV_CHECK_LAUNCH("center(k)");
  center_rows_kernel<<<1, 256, 0, s>>>(sp->to_v_w, sp->to_v_b, ps.Wv_c, ps.bv_c);
  SV_CHECK_LAUNCH("center(v)");
  transpose256_kernel<<<C, 256, 0, s>>>(ps.Wk_c, ps.Wk_cT);
  SV_CHECK_LAUNCH("transpose");
  SV_TRY(tc_prepare_stage(*sp, ps.Wk_c, ps.bk_c, ps.Wv_c, ps.bv_c, ps.tc, s));
  const int P = h * wd;
  SV_TRY(linear(slots_p, sp->to_q_w, sp->to_q_b, w.qraw, N, C, C, 0, nullptr, s));
  q_post_kernel<<<ceil_div(N, 8), 256, 0, s>>>(w.qraw, sp->nq_w, sp->nq_b, sp->nk_w, sp->nk_b, ps.bk_c, w.qt, w.g0, w.g1, N);
  SV_CHECK_LAUNCH("q_post");
  SV_TRY(linear_fast(w.qt, ps.Wk_cT, nullptr, w.G, N, C, C, 0, nullptr, s));
  const bool use_tc = kernel_path == 0 && tc_supported(&d, 0);
  if (use_tc) SV_TRY(tc_split_level(x, 0, pos, 0, false, w.tc, 1, h, wd, s));
  StagePix px;
  px.rs_k = w.rs_k; px.rs_v = w.rs_v; px.tc = w.tc;
  SV_TRY(pixel_attention(x, 0, pos, 0, ps, w, px, 1, N, P, use_tc, s));
  SV_TRY(linear(w.Z, ps.Wv_c, nullptr, w.Y, N, C, C, 0, nullptr, s));
  attn_post_kernel<<<ceil_div(N, 8), 256, 0, s>>>(w.Y, w.a0, w.a1, nullptr, sp->nv_w, sp->nv_b, ps.bv_c, sp->no_w, sp->no_b,
                                                   nullptr, nullptr, out, nullptr, N);
  SV_CHECK_LAUNCH("attn_post");
  return SLOTVPS_OK;
}

// ---- UPSNetFPN deformable-convolution subnet -------------------------------------------------
namespace {
int dcn_check_layers(const slotvps_dcn_layer* L, int n) {
  SV_REQUIRE(L != nullptr && n > 0 && n <= 8, "bad layer list");
  for (int i = 0; i < n; ++i) {
    SV_TRY(dcn::validate_layer(L[i]));
    SV_REQUIRE(i == 0 || L[i].c_in == L[i - 1].c_out, "dcn: c_in of a layer must equal c_out of the previous one");
  }
  return SLOTVPS_OK;
}
int dcn_cin_max(const slotvps_dcn_layer* L, int n) {
  int m = 0;
  for (int i = 0; i < n; ++i) m = L[i].c_in > m ? L[i].c_in : m;
  return m;
}
}  // namespace

int slotvps_dcn_prepared_bytes(const slotvps_dcn_layer* layers, int n_layers, size_t* bytes) {
  SV_REQUIRE(bytes != nullptr, "null out pointer");
  SV_TRY(dcn_check_layers(layers, n_layers));
  *bytes = align_up(dcn::prep_layout(layers, n_layers, nullptr, nullptr));
  return SLOTVPS_OK;
}

int slotvps_dcn_prepare(const slotvps_dcn_layer* layers, int n_layers, void* prepared, size_t prepared_bytes, void* stream) {
  SV_TRY(dcn_check_layers(layers, n_layers));
  SV_REQUIRE(prepared != nullptr, "null argument");
  if (prepared_bytes < align_up(dcn::prep_layout(layers, n_layers, nullptr, nullptr))) return fail(SLOTVPS_EWORKSPACE, "dcn: prepared buffer too small%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  dcn::LayerPrep lp[8];
  dcn::prep_layout(layers, n_layers, prepared, lp);
  for (int i = 0; i < n_layers; ++i) {
    SV_REQUIRE(layers[i].offset_w && layers[i].offset_b && layers[i].weight && layers[i].gn_w && layers[i].gn_b, "dcn: null parameter");
    const int cin = layers[i].c_in;
    dcn::offw_planes_kernel<<<ceil_div(C * cin, 256), 256, 0, s>>>(layers[i].offset_w, lp[i].offw, cin);
    SV_CHECK_LAUNCH("dcn_offw_prep");
    dcn::dcnw_prep_kernel<<<ceil_div(C * dcn::KT * cin, 256), 256, 0, s>>>(layers[i].weight, lp[i].wplanes, layers[i].c_out, cin);
    SV_CHECK_LAUNCH("dcn_w_prep");
  }
  return SLOTVPS_OK;
}

int slotvps_dcn_workspace_bytes(const slotvps_dcn_layer* layers, int n_layers, int B, int H, int W, size_t* bytes) {
  SV_REQUIRE(bytes != nullptr, "null out pointer");
  SV_TRY(dcn_check_layers(layers, n_layers));
  SV_REQUIRE(B > 0 && H > 0 && W > 0 && (long)B * H * W < (1L << 22), "bad shape (the kernels index pixels x channels in 32 bits: B*H*W < 2^22)");
  *bytes = align_up(dcn::ws_layout(dcn_cin_max(layers, n_layers), B, H, W, nullptr, nullptr));
  return SLOTVPS_OK;
}

int slotvps_dcn_subnet_forward(const slotvps_dcn_layer* layers, int n_layers, const void* prepared, const float* x, float* out,
                               int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  SV_TRY(dcn_check_layers(layers, n_layers));
  SV_REQUIRE(prepared && x && out && workspace, "null argument");
  SV_REQUIRE(B > 0 && H > 0 && W > 0 && (long)B * H * W < (1L << 22), "bad shape (the kernels index pixels x channels in 32 bits: B*H*W < 2^22)");
  const int cin_max = dcn_cin_max(layers, n_layers);
  if (workspace_bytes < align_up(dcn::ws_layout(cin_max, B, H, W, nullptr, nullptr))) return fail(SLOTVPS_EWORKSPACE, "dcn: workspace too small%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  dcn::LayerPrep lp[8];
  dcn::prep_layout(layers, n_layers, const_cast<void*>(prepared), lp);
  dcn::Ws w;
  dcn::ws_layout(cin_max, B, H, W, workspace, &w);
  const int P = H * W;
  const long rows = (long)B * P;
  {
    const int c0 = layers[0].c_in;
    dcn::nchw_to_nhwc_kernel<<<dim3(ceil_div(P, 32), ceil_div(c0, 32), B), 256, 0, s>>>(x, w.act, c0, P, w.aplanes, rows);   // + layer 0's operand planes
    SV_CHECK_LAUNCH("dcn_to_nhwc");
  }
  for (int i = 0; i < n_layers; ++i) {
    const int cin = layers[i].c_in, cout = layers[i].c_out;
    // the activation of this layer once: layer 0 = the input as it is; later = relu(GroupNorm(y)) of the previous layer, compacted to cin columns
    if (i > 0) {
      dcn::act_planes_kernel<<<dcn::grid_for(rows * (cin / 8)), 256, 0, s>>>(w.y, C, w.aff, w.act, w.aplanes, rows, cin, P);
      SV_CHECK_LAUNCH("dcn_act_planes");
    }
    // conv_offset: one 1x1 GEMM for all 9 taps x 18 outputs, then the 9-tap shift-sum
    SV_TRY(dcn::gemm(w.aplanes, rows, cin, lp[i].offw, w.z, 176, H, W, s));
    dcn::offset_shift_kernel<<<(unsigned)((rows * dcn::NOFF + 255) / 256), 256, 0, s>>>(w.z, layers[i].offset_b, w.off, H, W, rows);
    SV_CHECK_LAUNCH("dcn_offset_shift");
    dcn::Off off{w.off, (long)P * dcn::NOFF, dcn::NOFF, 1};
    // GroupNorm statistics: from the implicit-GEMM epilogue's per-tile partials when a tile lies in one image, else a pass over y
    const bool fused_stats = !dcn::use_im2col() && P % 128 == 0;
    SV_TRY(dcn::conv_gemm(w.act, off, lp[i].wplanes, w.planes, w.y, cin, cout, B, H, W, s, fused_stats ? w.tpart : nullptr));
    if (fused_stats) {
      dcn::gn_final_tiles_kernel<<<dim3(B, dcn::NG), 256, 0, s>>>(w.tpart, layers[i].gn_w, layers[i].gn_b, w.aff, P, cout);
    } else {
      dcn::gn_partial_kernel<<<dim3(w.slabs, B), 256, 0, s>>>(w.y, w.part, P, cout);
      SV_CHECK_LAUNCH("dcn_gn_partial");
      dcn::gn_final_kernel<<<B, 256, 0, s>>>(w.part, w.slabs, B, layers[i].gn_w, layers[i].gn_b, w.aff, P, cout);
    }
    SV_CHECK_LAUNCH("dcn_gn_final");
  }
  const int cout = layers[n_layers - 1].c_out;
  dcn::act_to_nchw_kernel<<<dim3(ceil_div(P, 32), ceil_div(cout, 32), B), 256, 0, s>>>(w.y, w.aff, out, P, cout);
  SV_CHECK_LAUNCH("dcn_to_nchw");
  return SLOTVPS_OK;
}

namespace {
struct DcWs { float *xT, *y; __half *wplanes, *planes; };
size_t dc_layout(int B, int c_in, int H, int W, void* base, DcWs* o) {
  Arena a(base, (size_t)-1);
  const size_t rows = (size_t)B * H * W;
  DcWs d;
  d.xT = a.take<float>(rows * c_in);
  d.y = a.take<float>(rows * C);
  d.wplanes = a.take<__half>((size_t)2 * C * dcn::KT * c_in);
  d.planes = a.take<__half>(dcn::use_im2col() ? (size_t)2 * rows * dcn::KT * c_in + 64 : 64);
  if (o) *o = d;
  return a.off;
}
}  // namespace

int slotvps_deform_conv_workspace_bytes(int B, int c_in, int H, int W, size_t* bytes) {
  SV_REQUIRE(bytes != nullptr, "null out pointer");
  SV_REQUIRE(B > 0 && H > 0 && W > 0 && c_in > 0 && c_in <= C && (long)B * H * W < (1L << 22), "bad shape (B*H*W < 2^22)");
  *bytes = align_up(dc_layout(B, c_in, H, W, nullptr, nullptr));
  return SLOTVPS_OK;
}

int slotvps_deform_conv_forward(const float* input, const float* weight, const float* offset, float* output, int B, int c_in, int c_out,
                                int H, int W, int kW, int kH, int dW, int dH, int padW, int padH, int dilationW, int dilationH,
                                int group, int deformable_group, int im2col_step, void* workspace, size_t workspace_bytes, void* stream) {
  (void)im2col_step;
  SV_REQUIRE(input && weight && output && workspace, "null argument");
  SV_REQUIRE(kW == 3 && kH == 3 && dW == 1 && dH == 1 && padW == 1 && padH == 1 && dilationW == 1 && dilationH == 1 && group == 1 &&
             deformable_group == 1, "deform_conv: only the UPSNetFPN instance (3x3, stride 1, padding 1, dilation 1, one group) is served");
  slotvps_dcn_layer l;
  memset(&l, 0, sizeof(l));
  l.c_in = c_in; l.c_out = c_out;
  SV_TRY(dcn::validate_layer(l));
  SV_REQUIRE(B > 0 && H > 0 && W > 0 && (long)B * H * W < (1L << 22), "bad shape (the kernels index pixels x channels in 32 bits: B*H*W < 2^22)");
  if (workspace_bytes < align_up(dc_layout(B, c_in, H, W, nullptr, nullptr))) return fail(SLOTVPS_EWORKSPACE, "deform_conv: workspace too small%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  SV_PROF_ENTRY();
  DcWs w;
  dc_layout(B, c_in, H, W, workspace, &w);
  const int P = H * W;
  dcn::dcnw_prep_kernel<<<ceil_div(C * dcn::KT * c_in, 256), 256, 0, s>>>(weight, w.wplanes, c_out, c_in);
  SV_CHECK_LAUNCH("dcn_w_prep");
  dcn::nchw_to_nhwc_kernel<<<dim3(ceil_div(P, 32), ceil_div(c_in, 32), B), 256, 0, s>>>(input, w.xT, c_in, P);
  SV_CHECK_LAUNCH("dcn_to_nhwc");
  dcn::Off off{offset, (long)P * dcn::NOFF, 1, (long)P};          // NCHW offsets: channel stride P
  SV_TRY(dcn::conv_gemm(w.xT, off, w.wplanes, w.planes, w.y, c_in, c_out, B, H, W, s));
  dcn::act_to_nchw_kernel<<<dim3(ceil_div(P, 32), ceil_div(c_out, 32), B), 256, 0, s>>>(w.y, nullptr, output, P, c_out);
  SV_CHECK_LAUNCH("dcn_to_nchw");
  return SLOTVPS_OK;
}

}  // extern "C"
