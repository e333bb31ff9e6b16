// Slot-update kernels, cluster form: the slot side of one retriever stage (dynamic_mask_head.py:342-400) with the
// frame's N <= 104 slots resident in a CLUSTER of four CTAs instead of one (slot_tc.cuh).
//
// Why: a 256 -> 256 linear layer on the 3-product fp16 hi/lo scheme is 6 K tensor-pipe cycles and 256 KB of weights for
// ONE SM, followed by a LayerNorm epilogue that cannot start before the last column is in; with seven such steps per
// stage the frame-resident kernels were a 50-90 us latency chain per launch on 2 of 148 SMs.  Here every CTA of the
// cluster owns 64 of the 256 output columns of every layer:
//   * its weight slice of a layer is 64 KB = the whole 8-slot TMA ring, so the next layer's weights are in shared
//     memory before the current epilogue ends (no refill round trips on the critical path);
//   * the MMAs per layer drop to 48 x (M128 N64 K16) = 1.5 K cycles;
//   * an epilogue thread owns ONE 16-column unit of its slot row, which stays in registers across the LayerNorm;
//   * row statistics go CTA-local through shared memory, then across the cluster through distributed shared memory;
//   * the normalised unit is written as fp16 hi/lo MMA operand into the CTA's own ACT planes and into an operand image in
//     global memory; after a cluster rendezvous every CTA bulk-copies the three foreign k-subtiles from L2, so each CTA
//     holds the full K = 256 operand of the next layer (distributed shared memory proved too slow for this: 17 B/clk).
// Cluster-wide synchronisation uses two alternating mbarriers per CTA (a CTA-local named barrier, then one release.cluster
// arrival per CTA on every CTA's barrier, acquire.cluster waits) because the TMA and MMA warps cannot take part in barrier.cluster.
//
//   slot_pre_cl     out_proj + residual + norm1 -> to_q + norm_q -> G = (q * gamma_k) Wk_c, g0, g1, fp16 planes of G
//   slot_post_cl    Wv_c Z, norm_v / norm1 / ReLU, residual, norm2 -> p2      (the FFN runs in slot_ffn_kernel)
//   slot_towers_cl  cls0 | reg0 -> LN/ReLU -> cls1 -> LN/ReLU -> logits ; reg0 -> LN/ReLU -> reg1 -> LN/ReLU -> next slots
#pragma once
#include "slot_tc.cuh"

namespace slotvps {
namespace slot {
namespace cl {
constexpr int CL = 4;                          // CTAs per frame
constexpr int NCOL = C / CL;                   // 64 output columns per CTA
constexpr int WT = NCOL * 128;                 // 8 KB weight tile [64 out][64 k] fp16, one plane
constexpr int NSL = 8;                         // ring slots = one layer slice (4 k-subtiles x hi/lo)
constexpr int OFF_RING = ACT_BYTES;
constexpr int OFF_STG = OFF_RING + NSL * WT;   // 16 per-warp staging tiles [32][20] floats
constexpr int OFF_RED = OFF_STG + EPI_WARPS * 2560;      // CTA-local partial sums [2][4][128] float2
constexpr int OFF_XCH = OFF_RED + 8192;        // cluster partial sums [2][4 ranks][128] float2
constexpr int OFF_BAR = OFF_XCH + 8192;
constexpr int OFF_RSC = OFF_BAR + 512;            // [128] floats: inverse row scales of the loaded rows
constexpr int SMEM = OFF_BAR + 1024 + 1024;
constexpr float INV_L = WSCALE_INV * LSCALE_INV;   // read-back scale of a GEMM on a LayerNorm-output operand
constexpr uint32_t IDESC64 = tc::make_idesc_f16(128, 64, 0, 0);
constexpr int CORR = 128;                      // TMEM column offset of the cross-product accumulators
static_assert(SMEM <= 232448, "shared memory budget");
static_assert(OFF_RING % 1024 == 0 && WT % 1024 == 0, "swizzle atoms");

struct Bars {
  uint64_t full[NSL], empty[NSL];
  uint64_t dfull, aready;
  uint64_t xbar[2];                             // cluster rendezvous of the epilogue warps
  uint64_t dq[3];                               // per-layer accumulator barriers where the issuer runs ahead (slot_tqkv_cl)
  uint64_t opfull;                              // the peers' operand slices have landed (bulk copies from the global image)
  uint32_t tmem_ptr;
};

__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cl_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void st_cl_v2f(uint32_t a, float x, float y) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAITC_DONE;\n\t"
      "bra WAITC_LOOP;\n\t"
      "WAITC_DONE:\n\t"
      "}" ::"r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> this CTA's shared memory, completion bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifdef SLOTVPS_SLOT_PROFILE
#define CL_MARK(i) tk[i] = clock64()
#else
#define CL_MARK(i)
#endif

// ---- TMA producer / MMA issuer: one 64-column slice of a layer -----------------------------------------------------
__device__ __forceinline__ void prod_slice(uint8_t* smem, Bars* b, uint32_t& it, const CUtensorMap* m, int hi_row, int lo_row) {
#pragma unroll 1
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll 1
    for (int pl = 0; pl < 2; ++pl) {
      const int s = it % NSL;
      tc::mbar_wait(&b->empty[s], ((it / NSL) & 1) ^ 1);
      tc::mbar_expect_tx(&b->full[s], WT);
      tc::tma_load_2d(smem + OFF_RING + s * WT, m, ks * 64, pl ? lo_row : hi_row, &b->full[s]);
      ++it;
    }
}
// The hi.hi products accumulate at d_tmem, the two cross products (2^-11 of the magnitude) at d_tmem + CORR: the fp32
// accumulator of the tensor pipe does not round to nearest, so every accumulation step costs up to an ulp of the running
// sum with a systematic sign; keeping the 32 small steps out of the main sum leaves 16 of the 48.
// Runs on ALL lanes of the MMA warp (warp-uniform waits and descriptors); `el` = tc::elect_one() guards the tcgen05 instructions.
__device__ __forceinline__ void mma_slice(bool el, uint8_t* smem, Bars* b, uint32_t& it, uint32_t d_tmem, uint32_t act, uint32_t idesc,
                                          uint32_t corr = CORR) {
  const uint32_t d_corr = d_tmem + corr;
#pragma unroll 1
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t dah = tc::make_smem_desc_sw128(act + ks * ACT_SUB, 16, 1024);
    const uint64_t dal = tc::make_smem_desc_sw128(act + ACT_PLANE + ks * ACT_SUB, 16, 1024);
    {
      const int s = it % NSL;
      tc::mbar_wait(&b->full[s], (it / NSL) & 1);
      tc::tc_fence_after();
      const uint64_t dbh = tc::make_smem_desc_sw128(tc::smem_u32(smem + OFF_RING + s * WT), 16, 1024);
      if (el) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc::umma_bf16(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, (ks != 0 || k != 0) ? 1u : 0u);
          tc::umma_bf16(d_corr, dal + 2 * k, dbh + 2 * k, idesc, (ks != 0 || k != 0) ? 1u : 0u);
        }
        tc::umma_commit(&b->empty[s]);
      }
      ++it;
    }
    {
      const int s = it % NSL;
      tc::mbar_wait(&b->full[s], (it / NSL) & 1);
      tc::tc_fence_after();
      const uint64_t dbl = tc::make_smem_desc_sw128(tc::smem_u32(smem + OFF_RING + s * WT), 16, 1024);
      if (el) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_bf16(d_corr, dah + 2 * k, dbl + 2 * k, idesc, 1);
        tc::umma_commit(&b->empty[s]);
      }
      ++it;
    }
  }
}

// ---- epilogue context --------------------------------------------------------------------------------------------
// 16 warps; warp e serves TMEM lane quadrant (e & 3) -- slot row 32 (e & 3) + lane -- and the 16-column unit (e >> 2) of the
// CTA's 64 columns, i.e. unit u = 4 rank + (e >> 2) of the layer's 256.
struct Ctx {
  Bars* b;
  Epi e;                                        // row / lane bookkeeping and the staging tile (read_rows / write_rows)
  int u;                                        // this thread's unit of the 256 columns
  uint32_t rank, nred, nx, nsync, nd;
  float2 *red, *xch;
  uint32_t xch_r[CL];                           // shared::cluster addresses of every CTA's XCH base
  uint8_t *smem, *gx;                           // this CTA's shared memory; the frame's operand image in global memory [2][4][104][128 B]
  // rendezvous of the epilogue warps of the whole cluster; also orders this CTA's earlier shared::cluster / global stores
  // before the peers' reads.  CTA-local named barrier first, then ONE thread arrives on the barrier of every CTA: with one
  // arrival per warp (64 remote atomics on each barrier) a rendezvous cost 2-2.5 K cycles.
  __device__ __forceinline__ void xsync() {
    e.sync();
    uint64_t* bar = &b->xbar[nsync & 1];
    if (threadIdx.x == 64) {
#pragma unroll
      for (uint32_t d = 0; d < CL; ++d) tc::mbar_arrive_remote(bar, d);
    }
    mbar_wait_cl(bar, (nsync >> 1) & 1);
    ++nsync;
  }
  __device__ __forceinline__ void wait_d() { tc::mbar_wait(&b->dfull, nd & 1); ++nd; tc::tc_fence_after(); }
  __device__ __forceinline__ void publish() { tc::tc_fence_before(); tc::fence_proxy_async(); tc::mbar_arrive(&b->aready); }
  // accumulator columns [col0 + 16 qt, +16) of this thread's row (main + cross products) times the inverse operand scales
  __device__ __forceinline__ void ld(int col0, float* v, float inv, int corr = CORR) const {
    float cr[16];
    tc::tmem_ld16(e.tbase + col0 + 16 * e.qt, v);
    tc::tmem_ld16(e.tbase + corr + col0 + 16 * e.qt, cr);
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = (v[c] + cr[c]) * inv;
  }
  // (a, q) of this thread's 16 columns -> per-CTA pairs of the four CTAs, in rank order.  CTA-local combine `comb` folds
  // the four column quarters of a row into one pair; the rendezvous also orders earlier st.shared::cluster of this CTA.
  template <class Comb>
  __device__ __forceinline__ void row_gather(float a, float q, Comb&& comb, float2* out) {
    float2* rb = red + (nred & 1) * 512;
    ++nred;
    rb[e.qt * 128 + e.r] = make_float2(a, q);
    e.sync();
    const uint32_t slot = (nx & 1) * 512;
    ++nx;
    if (e.qt == 0) {
      const float2 s = comb(rb[e.r], rb[128 + e.r], rb[256 + e.r], rb[384 + e.r]);
      const uint32_t off = (slot + rank * 128 + e.r) * 8;
#pragma unroll
      for (int d = 0; d < CL; ++d) st_cl_v2f(xch_r[d] + off, s.x, s.y);
    }
    xsync();
    const float2* xb = xch + slot;
    out[0] = xb[e.r]; out[1] = xb[128 + e.r]; out[2] = xb[256 + e.r]; out[3] = xb[384 + e.r];
  }
  // plain sums of two per-thread values over the row's 256 columns
  __device__ __forceinline__ void row_total(float a, float q, float& ta, float& tq) {
    float2 x[4];
    row_gather(a, q, [](float2 p0, float2 p1, float2 p2, float2 p3) {
      return make_float2((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y)); }, x);
    ta = (x[0].x + x[1].x) + (x[2].x + x[3].x);
    tq = (x[0].y + x[1].y) + (x[2].y + x[3].y);
  }
  // LayerNorm of the row whose unit is v, optional ReLU.  Mean and CENTRED second moment are combined pairwise
  // (Chan et al.): per thread over 16 values, per CTA over 4 threads, per row over 4 CTAs -- as accurate as the two-pass
  // form of the oracle with a single exchange.
  __device__ __forceinline__ void layer_norm(float* v, const float* __restrict__ gw, const float* __restrict__ gb, bool relu) {
    float sp = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) sp += v[c];
    const float m16 = sp * (1.f / 16.f);
    float m2 = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { const float dl = v[c] - m16; m2 = fmaf(dl, dl, m2); }
    float2 x[4];
    row_gather(m16, m2, [](float2 p0, float2 p1, float2 p2, float2 p3) {        // 4 x (mean, M2) of 16 -> (mean, M2) of 64
      const float m = 0.25f * ((p0.x + p1.x) + (p2.x + p3.x));
      const float d0 = p0.x - m, d1 = p1.x - m, d2 = p2.x - m, d3 = p3.x - m;
      return make_float2(m, ((p0.y + p1.y) + (p2.y + p3.y)) + 16.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3))); }, x);
    const float mean = 0.25f * ((x[0].x + x[1].x) + (x[2].x + x[3].x));
    const float d0 = x[0].x - mean, d1 = x[1].x - mean, d2 = x[2].x - mean, d3 = x[3].x - mean;
    const float M2 = ((x[0].y + x[1].y) + (x[2].y + x[3].y)) + 64.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    const float rstd = rsqrtf(M2 * (1.f / C) + LN_EPS);
    float w[16], bb[16];
    ldg16(gw + 16 * u, w); ldg16(gb + 16 * u, bb);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      v[c] = (v[c] - mean) * rstd * w[c] + bb[c];
      if (relu) v[c] = fmaxf(v[c], 0.f);
    }
  }
  // LayerNorm-output unit (times LSCALE) -> fp16 hi/lo operand planes: this CTA's k-subtile (= its rank) in its own shared
  // memory and in the frame's operand image in global memory, from where the peers fetch it (publish_all).  Distributed
  // shared memory moves ~17 B/clk per SM -- writing the slice into four CTAs took 8.3 K cycles per layer -- while the L2
  // path delivers the three foreign slices (80 KB) at the ~64 B/clk ingest rate of an SM.
  __device__ __forceinline__ void to_act_all(const float* v) const {
    if (e.r >= NR) return;
    const uint32_t off = rank * ACT_SUB + e.r * 128;               // u >> 2 == rank
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
      if (e.valid) {
#pragma unroll
        for (int k = 0; k < 4; ++k) split2(v[8 * q + 2 * k] * LSCALE, v[8 * q + 2 * k + 1] * LSCALE, hi[k], lo[k]);
      }
      const uint32_t phys = off + ((((u & 3) * 2 + q) ^ (e.r & 7))) * 16;
      const uint4 h4 = make_uint4(hi[0], hi[1], hi[2], hi[3]), l4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(smem + OFF_ACT + phys) = h4;
      *reinterpret_cast<uint4*>(smem + OFF_ACT + ACT_PLANE + phys) = l4;
      *reinterpret_cast<uint4*>(gx + phys) = h4;
      *reinterpret_cast<uint4*>(gx + ACT_PLANE + phys) = l4;
    }
  }
  // own slice written -> rendezvous -> fetch the peers' slices; the MMA issuer waits for aready and opfull
  __device__ __forceinline__ void publish_fetch() {
    if (threadIdx.x == 64) {
      fence_proxy_async_all();
      tc::mbar_expect_tx(&b->opfull, 6 * ACT_SUB);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
          if (ks != rank) bulk_g2s(smem + OFF_ACT + pl * ACT_PLANE + ks * ACT_SUB, gx + pl * ACT_PLANE + ks * ACT_SUB, ACT_SUB, &b->opfull);
    }
    tc::tc_fence_before();
    tc::mbar_arrive(&b->aready);
  }
  __device__ __forceinline__ void publish_all() {
    fence_proxy_async_all();                                       // generic writes (shared and global) before async-proxy reads
    xsync();
    publish_fetch();
  }
};

__device__ __forceinline__ Ctx make_ctx(uint8_t* smem, Bars* b, uint32_t tmem_base, uint32_t rank, int N, uint8_t* gx) {
  Ctx c;
  const int we = (threadIdx.x >> 5) - 2;
  c.b = b; c.rank = rank; c.nred = 0; c.nx = 0; c.nsync = 0; c.nd = 0;
  c.e.lane = threadIdx.x & 31; c.e.q = (threadIdx.x >> 5) & 3; c.e.qt = we >> 2; c.e.r = c.e.q * 32 + c.e.lane; c.e.N = N;
  c.e.valid = c.e.r < N;
  c.e.tbase = tmem_base + ((uint32_t)(c.e.q * 32) << 16);
  c.e.stg = reinterpret_cast<float*>(smem + OFF_STG) + we * 640;
  c.u = (int)rank * 4 + c.e.qt;
  c.red = reinterpret_cast<float2*>(smem + OFF_RED);
  c.xch = reinterpret_cast<float2*>(smem + OFF_XCH);
#pragma unroll
  for (uint32_t d = 0; d < CL; ++d) c.xch_r[d] = mapa(tc::smem_u32(smem + OFF_XCH), d);
  c.smem = smem; c.gx = gx;
  return c;
}

// barriers, TMEM, cluster rendezvous; returns the TMEM base
__device__ __forceinline__ uint32_t prologue(Bars* b, int warp, uint32_t tmem_cols) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSL; ++i) { tc::mbar_init(&b->full[i], 1); tc::mbar_init(&b->empty[i], 1); }
    tc::mbar_init(&b->dfull, 1); tc::mbar_init(&b->aready, EPI_THREADS);
    tc::mbar_init(&b->xbar[0], CL); tc::mbar_init(&b->xbar[1], CL);
    for (int i = 0; i < 3; ++i) tc::mbar_init(&b->dq[i], 1);
    tc::mbar_init(&b->opfull, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc(&b->tmem_ptr, tmem_cols); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();                                               // every CTA's barriers exist before a peer targets them
  tc::tc_fence_after();
  return __shfl_sync(0xffffffffu, b->tmem_ptr, 0);                  // provably warp-uniform (uniform-register descriptors)
}
__device__ __forceinline__ void epilogue_exit(uint32_t tmem_base, int warp, uint32_t tmem_cols) {
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();                                               // no CTA leaves while a peer may still write to it
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, tmem_cols); }
}

// =====================================================================================================================
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
slot_pre_cl(const __grid_constant__ CUtensorMap m_out, const __grid_constant__ CUtensorMap m_q, const __grid_constant__ CUtensorMap m_wk,
            const PreParams P) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  Bars* b = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, vf = blockIdx.x / CL;
  const RowGroup rg = row_group(vf, P.N, P.G);
  const int t = rg.t, N = rg.n;                 // frame; slot rows of this cluster's group
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) { tc::tma_prefetch_desc(&m_out); tc::tma_prefetch_desc(&m_q); tc::tma_prefetch_desc(&m_wk); }
  const uint32_t tmem_base = prologue(b, warp, 256);
  const int hi_row = (int)rank * NCOL, lo_row = C + (int)rank * NCOL;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      prod_slice(smem, b, it, &m_out, hi_row, lo_row);
      prod_slice(smem, b, it, &m_q, hi_row, lo_row);
      prod_slice(smem, b, it, &m_wk, hi_row, lo_row);
    }
  } else if (warp == 1) {
    {                                                             // all lanes: warp-uniform issue loop (see tc::elect_one)
      const bool el = tc::elect_one();
      uint32_t it = 0;
      const uint32_t act = tc::smem_u32(smem + OFF_ACT);
      for (int g = 0; g < 3; ++g) {
        tc::mbar_wait(&b->aready, g & 1);
        if (g > 0) tc::mbar_wait(&b->opfull, (g - 1) & 1);        // the peers' slices of the LayerNorm-output operand
        tc::tc_fence_after();
        mma_slice(el, smem, b, it, tmem_base, act, IDESC64);
        if (el) tc::umma_commit(&b->dfull);
      }
    }
  } else {
    Ctx c = make_ctx(smem, b, tmem_base, rank, N, P.opx + (long)vf * ACT_BYTES);
    Epi& e = c.e;
    const int u = c.u;
    const long fbase = rg.rb * C;
#ifdef SLOTVPS_SLOT_PROFILE
    long long tk[20];
#endif
    CL_MARK(0);
    // A operand of the first GEMM: the self-attention output rows (every CTA builds the full K = 256 operand)
    float* rsc = reinterpret_cast<float*>(smem + OFF_RSC);
    load_rows_to_act(smem, P.mo + fbase, N, warp - 2, lane, rsc);
    c.publish();
    CL_MARK(1);
    float v[16], add[16];
    // ---- (1) out_proj + residual + norm1 -> p ----
    {
      float bb[16];
      e.read_rows(P.slots + fbase, u, add);
      ldg16(P.out_b + 16 * u, bb);
#pragma unroll
      for (int k = 0; k < 16; ++k) add[k] += bb[k];
      prefetch_l1(P.n1_w + 16 * u); prefetch_l1(P.n1_b + 16 * u);
    }
    e.sync();                                                      // rsc is complete
    const float inv0 = WSCALE_INV * rsc[e.r];
    CL_MARK(2);
    c.wait_d();
    CL_MARK(3);
    c.ld(0, v, inv0);
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] += add[k];
    c.layer_norm(v, P.n1_w, P.n1_b, false);
    CL_MARK(4);
    c.to_act_all(v);
    e.write_rows(P.p + fbase, u, v);
    ldg16(P.q_b + 16 * u, add);
    prefetch_l1(P.nq_w + 16 * u); prefetch_l1(P.nq_b + 16 * u);
    prefetch_l1(P.nk_w + 16 * u); prefetch_l1(P.nk_b + 16 * u); prefetch_l1(P.bk_c + 16 * u);
    CL_MARK(5);
    c.publish_all();
    CL_MARK(6);
    // ---- (2) to_q + norm_q -> q; qt = q * gamma_k; g0 = qt . bk_c; g1 = q . beta_k ----
    c.wait_d();
    CL_MARK(7);
    c.ld(0, v, INV_L);
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] += add[k];
    CL_MARK(11);
    c.layer_norm(v, P.nq_w, P.nq_b, false);
    CL_MARK(12);
    float s0 = 0.f, s1 = 0.f;
    {
      float gk[16], bk[16], bc[16];
      ldg16(P.nk_w + 16 * u, gk); ldg16(P.nk_b + 16 * u, bk); ldg16(P.bk_c + 16 * u, bc);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float qv = v[k], tv = qv * gk[k];
        s0 = fmaf(tv, bc[k], s0);
        s1 = fmaf(qv, bk[k], s1);
        v[k] = tv;
      }
    }
    CL_MARK(13);
    c.to_act_all(v);
    CL_MARK(14);
    c.row_total(s0, s1, s0, s1);
    if (e.valid && e.qt == 0 && rank == 0) { P.g0[rg.rb + e.r] = s0; P.g1[rg.rb + e.r] = s1; }
    CL_MARK(15);
    fence_proxy_async_all();
    CL_MARK(16);
    c.xsync();
    CL_MARK(17);
    c.publish_fetch();
    CL_MARK(8);
    // ---- (3) G = qt . Wk_c -> fp32 + fp16 hi/lo planes (the B operand of attn_tc's S product) ----
    c.wait_d();
    CL_MARK(9);
    c.ld(0, v, INV_L);
    if (P.Gout) e.write_rows(P.Gout + fbase, u, v);
    if (P.gplanes) {
      uint32_t* stw = reinterpret_cast<uint32_t*>(e.stg);           // staging as [2 planes][32 rows][10 words] (8 used)
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        hi[k] = 0u; lo[k] = 0u;
        if (e.valid) split2(v[2 * k], v[2 * k + 1], hi[k], lo[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        *reinterpret_cast<uint2*>(stw + lane * 10 + k) = make_uint2(hi[k], hi[k + 1]);
        *reinterpret_cast<uint2*>(stw + 320 + lane * 10 + k) = make_uint2(lo[k], lo[k + 1]);
      }
      __syncwarp();
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int it = 0; it < 4; ++it) {                            // a row's 16 columns = 32 bytes = 4 lanes x 8 bytes; 8 rows per request
          const int row = it * 8 + (lane >> 2), gr = e.q * 32 + row, part = lane & 3;
          if (gr < NR) {
            const uint2 w2 = *reinterpret_cast<const uint2*>(stw + pl * 320 + row * 10 + part * 2);
            __half* dst = P.gplanes + (((long)t * 2 + pl) * NR + gr) * C + 16 * u + part * 4;
            *reinterpret_cast<uint2*>(dst) = w2;
          }
        }
      __syncwarp();
    }
    CL_MARK(10);
#ifdef SLOTVPS_SLOT_PROFILE
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 64)
      printf("slot_pre_cl cycles: load %lld | prefetch %lld | wait1 %lld | ld+LN1 %lld | emit1 %lld | sync O %lld | wait2 %lld | epi2 %lld | wait3 %lld | epi3 %lld\n",
             tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], tk[7] - tk[6], tk[8] - tk[7], tk[9] - tk[8], tk[10] - tk[9]);
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 64)
      printf("   epi2: ld %lld | layer_norm %lld | gk math %lld | to_act %lld | row_total %lld | fence %lld | xsync %lld | fetch+arrive %lld\n",
             tk[11] - tk[7], tk[12] - tk[11], tk[13] - tk[12], tk[14] - tk[13], tk[15] - tk[14], tk[16] - tk[15], tk[17] - tk[16], tk[8] - tk[17]);
#endif
  }
  epilogue_exit(tmem_base, warp, 256);
}

// =====================================================================================================================
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
slot_post_cl(const __grid_constant__ CUtensorMap m_wv, const PostParams P) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  Bars* b = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, vf = blockIdx.x / CL;
  const RowGroup rg = row_group(vf, P.N, P.G);
  const int t = rg.t, N = rg.n;                 // frame; slot rows of this cluster's group
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) tc::tma_prefetch_desc(&m_wv);
  const uint32_t tmem_base = prologue(b, warp, 256);
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      prod_slice(smem, b, it, &m_wv, (int)rank * NCOL, C + (int)rank * NCOL);
    }
  } else if (warp == 1) {
    {                                                             // all lanes: warp-uniform issue loop (see tc::elect_one)
      const bool el = tc::elect_one();
      uint32_t it = 0;
      tc::mbar_wait(&b->aready, 0);
      tc::tc_fence_after();
      mma_slice(el, smem, b, it, tmem_base, tc::smem_u32(smem + OFF_ACT), IDESC64);       // Y = Z . Wv_c^T
      if (el) tc::umma_commit(&b->dfull);
    }
  } else {
    Ctx c = make_ctx(smem, b, tmem_base, rank, N, nullptr);
    Epi& e = c.e;
    const int u = c.u;
    const long fbase = rg.rb * C;
    float* rsc = reinterpret_cast<float*>(smem + OFF_RSC);
    load_rows_to_act(smem, P.Z + fbase, N, warp - 2, lane, rsc);
    c.publish();
    // ---- value projection of the pixel-reduced slots, norm_v / norm1 / ReLU, residual, norm2 (:456-459, 374-376) ----
    float v[16], pp[16];
    e.read_rows(P.p + fbase, u, pp);
    const float a0r = e.valid ? P.a0[rg.rb + e.r] : 0.f, a1r = e.valid ? P.a1[rg.rb + e.r] : 0.f;
    prefetch_l1(P.nv_w + 16 * u); prefetch_l1(P.nv_b + 16 * u); prefetch_l1(P.bv_c + 16 * u);
    prefetch_l1(P.no_w + 16 * u); prefetch_l1(P.no_b + 16 * u); prefetch_l1(P.n2_w + 16 * u); prefetch_l1(P.n2_b + 16 * u);
    e.sync();                                                      // rsc is complete
    const float inv0 = WSCALE_INV * rsc[e.r];
    c.wait_d();
    c.ld(0, v, inv0);
    {
      float gv[16], bv[16], bc[16];
      ldg16(P.nv_w + 16 * u, gv); ldg16(P.nv_b + 16 * u, bv); ldg16(P.bv_c + 16 * u, bc);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = gv[k] * fmaf(bc[k], a1r, v[k]) + bv[k] * a0r;
    }
    c.layer_norm(v, P.no_w, P.no_b, true);
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] += pp[k];
    c.layer_norm(v, P.n2_w, P.n2_b, false);
    e.write_rows(P.p2buf + fbase, u, v);
  }
  epilogue_exit(tmem_base, warp, 256);
}

// =====================================================================================================================
// towers (:390-400).  Accumulator columns [0, 64) = cls chain, [64, 128) = first reg layer (kept until the cls chain is done).
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
slot_towers_cl(const __grid_constant__ CUtensorMap m_tw, const __grid_constant__ CUtensorMap m_c1, const __grid_constant__ CUtensorMap m_lg,
               const __grid_constant__ CUtensorMap m_r1, const PostParams P) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  Bars* b = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, vf = blockIdx.x / CL;
  const RowGroup rg = row_group(vf, P.N, P.G);
  const int t = rg.t, N = rg.n;                 // frame; slot rows of this cluster's group
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) { tc::tma_prefetch_desc(&m_tw); tc::tma_prefetch_desc(&m_c1); tc::tma_prefetch_desc(&m_lg); tc::tma_prefetch_desc(&m_r1); }
  const uint32_t tmem_base = prologue(b, warp, 256);
  const int hi_row = (int)rank * NCOL, lo_row = C + (int)rank * NCOL;
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      prod_slice(smem, b, it, &m_tw, hi_row, 2 * C + hi_row);                     // cls0 slice  (planes [2][512][256])
      prod_slice(smem, b, it, &m_tw, C + hi_row, 3 * C + hi_row);                 // reg0 slice
      prod_slice(smem, b, it, &m_c1, hi_row, lo_row);
      prod_slice(smem, b, it, &m_lg, 0, TILE_N);                                  // class logits: 32 padded rows used of the 64-row box
      prod_slice(smem, b, it, &m_r1, hi_row, lo_row);
    }
  } else if (warp == 1) {
    {                                                             // all lanes: warp-uniform issue loop (see tc::elect_one)
      const bool el = tc::elect_one();
      uint32_t it = 0, na = 0;
      const uint32_t act = tc::smem_u32(smem + OFF_ACT);
      auto wait_act = [&]() {                                       // after the first: also the peers' operand slices
        tc::mbar_wait(&b->aready, na & 1);
        if (na > 0) tc::mbar_wait(&b->opfull, (na - 1) & 1);
        ++na;
        tc::tc_fence_after();
      };
      wait_act();
      mma_slice(el, smem, b, it, tmem_base, act, IDESC64);                            // cls0
      mma_slice(el, smem, b, it, tmem_base + 64, act, IDESC64);                       // reg0
      if (el) tc::umma_commit(&b->dfull);
      wait_act();
      mma_slice(el, smem, b, it, tmem_base, act, IDESC64);                            // cls1
      if (el) tc::umma_commit(&b->dfull);
      wait_act();
      mma_slice(el, smem, b, it, tmem_base, act, IDESC32);                            // class logits (every CTA; rank 0 stores)
      if (el) tc::umma_commit(&b->dfull);
      wait_act();
      mma_slice(el, smem, b, it, tmem_base, act, IDESC64);                            // reg1
      if (el) tc::umma_commit(&b->dfull);
    }
  } else {
    Ctx c = make_ctx(smem, b, tmem_base, rank, N, P.opx + (long)vf * ACT_BYTES);
    Epi& e = c.e;
    const int u = c.u;
    const long fbase = rg.rb * C;
    float* rsc = reinterpret_cast<float*>(smem + OFF_RSC);
    load_rows_to_act(smem, P.f_in + fbase, N, warp - 2, lane, rsc);
    c.publish();
    prefetch_l1(P.tw_ln_w + 16 * u); prefetch_l1(P.tw_ln_b + 16 * u); prefetch_l1(P.tw_ln_w + C + 16 * u); prefetch_l1(P.tw_ln_b + C + 16 * u);
    prefetch_l1(P.c1_nw + 16 * u); prefetch_l1(P.c1_nb + 16 * u); prefetch_l1(P.r1_nw + 16 * u); prefetch_l1(P.r1_nb + 16 * u);
    float v[16];
    e.sync();                                                      // rsc is complete
    const float inv0 = WSCALE_INV * rsc[e.r];
    c.wait_d();
    c.ld(0, v, inv0);
    c.layer_norm(v, P.tw_ln_w, P.tw_ln_b, true);                    // c1
    c.to_act_all(v);
    c.publish_all();
    c.wait_d();
    c.ld(0, v, INV_L);
    c.layer_norm(v, P.c1_nw, P.c1_nb, true);                        // c2
    c.to_act_all(v);
    c.publish_all();
    c.wait_d();
    if (rank == 0 && e.qt == 0) {
      float lg[2][16], lc[2][16];
      tc::tmem_ld16(e.tbase, lg[0]); tc::tmem_ld16(e.tbase + 16, lg[1]);
      tc::tmem_ld16(e.tbase + CORR, lc[0]); tc::tmem_ld16(e.tbase + CORR + 16, lc[1]);
      tc::tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 16; ++k) { lg[0][k] += lc[0][k]; lg[1][k] += lc[1][k]; }
      if (e.valid) {
        float* dst = P.cls_out + (long)t * P.cls_fs + (long)(rg.row0 + e.r) * P.ncls;
        for (int k = 0; k < P.ncls; ++k) dst[k] = lg[k >> 4][k & 15] * INV_L + __ldg(P.logit_b + k);
      }
    }
    c.ld(64, v, inv0);                                              // first reg layer, parked since the first GEMM
    c.layer_norm(v, P.tw_ln_w + C, P.tw_ln_b + C, true);            // e1
    c.to_act_all(v);
    c.publish_all();
    c.wait_d();
    c.ld(0, v, INV_L);
    c.layer_norm(v, P.r1_nw, P.r1_nb, true);                        // next-stage slots = this stage's embedding
    e.write_rows(P.slots_out + fbase, u, v, P.emb_out + (long)t * P.emb_fs + (long)rg.row0 * C);
  }
  epilogue_exit(tmem_base, warp, 256);
}

// =====================================================================================================================
// Video Retriever projections (:494-527): q | k | v = LN_j(f W_j^T + b_j) for the frame's slots -> tqkv rows r*3 + j.
// The three layers share the operand, so their MMAs run back to back and no operand exchange is needed.
struct TqkvParams {
  int N, G;                                     // slots per frame, row groups per frame
  const float* f_in;                            // [T][N][256]
  const float *bias, *ln_w, *ln_b;              // [3][256] each
  float* tqkv;                                  // [T*N][3][256]
};
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
slot_tqkv_cl(const __grid_constant__ CUtensorMap m_qkv, const TqkvParams P) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  Bars* b = reinterpret_cast<Bars*>(smem + OFF_BAR);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, vf = blockIdx.x / CL;
  const RowGroup rg = row_group(vf, P.N, P.G);
  const int t = rg.t, N = rg.n;                 // frame; slot rows of this cluster's group
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) tc::tma_prefetch_desc(&m_qkv);
  const uint32_t tmem_base = prologue(b, warp, 512);
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int j = 0; j < 3; ++j) prod_slice(smem, b, it, &m_qkv, j * C + (int)rank * NCOL, 3 * C + j * C + (int)rank * NCOL);
    }
  } else if (warp == 1) {
    {                                                             // all lanes: warp-uniform issue loop (see tc::elect_one)
      const bool el = tc::elect_one();
      uint32_t it = 0;
      tc::mbar_wait(&b->aready, 0);
      tc::tc_fence_after();
      for (int j = 0; j < 3; ++j) {
        mma_slice(el, smem, b, it, tmem_base + j * NCOL, tc::smem_u32(smem + OFF_ACT), IDESC64, 256);
        if (el) tc::umma_commit(&b->dq[j]);
      }
    }
  } else {
    Ctx c = make_ctx(smem, b, tmem_base, rank, N, nullptr);
    Epi& e = c.e;
    const int u = c.u;
    float* rsc = reinterpret_cast<float*>(smem + OFF_RSC);
    load_rows_to_act(smem, P.f_in + rg.rb * C, N, warp - 2, lane, rsc);
    c.publish();
    for (int j = 0; j < 3; ++j) { prefetch_l1(P.bias + j * C + 16 * u); prefetch_l1(P.ln_w + j * C + 16 * u); prefetch_l1(P.ln_b + j * C + 16 * u); }
    e.sync();                                                      // rsc is complete
    const float inv0 = WSCALE_INV * rsc[e.r];
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      float v[16], bb[16];
      ldg16(P.bias + j * C + 16 * u, bb);
      tc::mbar_wait(&b->dq[j], 0);
      tc::tc_fence_after();
      c.ld(j * NCOL, v, inv0, 256);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] += bb[k];
      c.layer_norm(v, P.ln_w + j * C, P.ln_b + j * C, false);
      e.write_rows(P.tqkv + (rg.rb * 3 + j) * C, u, v, nullptr, 3 * C);
    }
  }
  epilogue_exit(tmem_base, warp, 512);
}

// tensor map over [2 * Opad][K] fp16 planes, box [64][64]
inline int slot_wmap64(CUtensorMap* m, const __half* planes, int O, int K) {
  const int Opad = ceil_div(O, slot::TILE_N) * slot::TILE_N;
  return tc::make_tmap_h16_sw128(m, planes, (uint64_t)2 * Opad, (uint64_t)K, NCOL);
}

}  // namespace cl
}  // namespace slot
}  // namespace slotvps
