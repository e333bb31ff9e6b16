// Slot-update kernels on the tensor pipe (north_star kernel 2): the slot side of one retriever stage
// (dynamic_mask_head.py:342-400) with the N <= 104 slots of a frame RESIDENT in one CTA -- activations live in
// shared memory as fp16 hi/lo MMA operands and in TMEM as fp32 accumulators, the stage's weights stream through
// a TMA ring, every LayerNorm / GELU / residual runs thread-per-slot-row straight out of TMEM.
//
//   slot_pre_kernel   (after the slot self-attention core):  out_proj + residual + norm1  ->  to_q + norm_q
//                     -> folded key operands  G = (q*gamma_k) Wk_c, g0, g1  -> fp16 hi/lo planes for attn_tc
//   slot_post_kernel  (after the pixel attention):  Wv_c Z, norm_v / norm1 / ReLU / residual / norm2  ->  FFN
//                     256 -> F -> 256 in 128-wide hidden chunks (the hidden activation never leaves the SM)  -> norm3
//                     [-> Video Retriever on the generic kernels] -> cls / reg towers -> class logits, next-stage slots
//
// One CTA per frame (grid = T), 192 threads: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (thread r owns slot row r = TMEM lane r).  GEMM shape: M = 128 slot rows (lanes),
// N = 128 output features per weight tile, K = 64 per ring slot; fp16 hi/lo operands, 3 products, fp32 accumulation.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "pixel_tc.cuh"

namespace slotvps {
namespace slot {
constexpr int NR = attn::NROW;                 // 104 slot rows held per CTA
constexpr int TILE_N = 128;                    // output features per weight tile
constexpr int W_TILE = TILE_N * 128;           // 16 KB  [128 out][64 k] fp16, one plane
constexpr int NSLOT = 4;                       // ring of weight tiles (hi, lo, hi, lo ...)
constexpr int ACT_SUB = NR * 128;              // 13312 B [104 rows][64 k] fp16
constexpr int ACT_PLANE = 4 * ACT_SUB;         // K = 256
constexpr int ACT_BYTES = 2 * ACT_PLANE;       // hi + lo = 106496
constexpr int HB_PLANE = 2 * ACT_SUB;          // hidden chunk: K = 128
constexpr int HB_BYTES = 2 * HB_PLANE;         // 53248
constexpr int OFF_ACT = 0;
constexpr int OFF_HB = OFF_ACT + ACT_BYTES;
constexpr int OFF_RING = OFF_HB + HB_BYTES;
constexpr int OFF_MISC = OFF_RING + NSLOT * W_TILE;
constexpr int MISC_BYTES = 1024;
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int THREADS = 192;
constexpr uint32_t IDESC128 = tc::make_idesc_f16(128, 128, 0, 0);
constexpr uint32_t IDESC32 = tc::make_idesc_f16(128, 32, 0, 0);
constexpr int TM_D = 0;                        // main accumulator, 256 columns (towers: 512)
constexpr int TM_D1 = 256;                     // two 128-column FFN hidden accumulators
// Weight planes are stored scaled by 2^8: fp16 has 5 exponent bits, so the lo plane of a weight of ~0.05 would be
// subnormal (absolute step 6e-8, ~1e-6 relative instead of 2.4e-7 -- and for sums of random-sign terms the relative
// error of the sum equals the per-term one).  Every accumulator read is multiplied by 2^-8 (exact).
constexpr float WSCALE = 256.f, WSCALE_INV = 1.f / 256.f;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(OFF_HB % 1024 == 0 && OFF_RING % 1024 == 0 && ACT_SUB % 1024 == 0, "swizzle atoms");

// fp16 hi/lo weight planes of one nn.Linear for the TMA ring: [2][Opad][K], Opad = ceil(O / 128) * 128 (zero rows)
__global__ void __launch_bounds__(256) linear_planes_kernel(const float* __restrict__ W, int O, int K, int Opad, __half* __restrict__ out) {
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long)Opad * K) return;
  const int o = (int)(i / K);
  __half h = __float2half_rn(0.f), l = h;
  if (o < O) split_bf16(W[i] * WSCALE, h, l);
  out[i] = h;
  out[(long)Opad * K + i] = l;
}

struct Barriers {
  uint64_t full[NSLOT], empty[NSLOT];
  uint64_t dfull, aready;                      // accumulator complete (MMA -> epilogue) / operand written (epilogue -> MMA)
  uint64_t d1full[2], d1free[2], hfull, hfree; // FFN pipeline
  uint32_t tmem_ptr;
};

// ---- device helpers shared by both kernels --------------------------------------------------------------------
struct Ring {                                   // producer / issuer side bookkeeping of the weight ring
  uint32_t it = 0;
};

__device__ __forceinline__ void prod_tile(uint8_t* smem, Barriers* b, Ring& rg, const CUtensorMap* m, int k0, int row) {
  const int s = rg.it % NSLOT;
  tc::mbar_wait(&b->empty[s], ((rg.it / NSLOT) & 1) ^ 1);
  tc::mbar_expect_tx(&b->full[s], W_TILE);
  tc::tma_load_2d(smem + OFF_RING + s * W_TILE, m, k0, row, &b->full[s]);
  ++rg.it;
}
// weights of `ntiles` output tiles x `nks` k-subtiles, hi then lo plane per (tile, k-subtile)
__device__ __forceinline__ void prod_gemm(uint8_t* smem, Barriers* b, Ring& rg, const CUtensorMap* m, int opad, int n_first, int ntiles,
                                          int ks_first, int nks) {
  for (int nt = 0; nt < ntiles; ++nt)
    for (int ks = 0; ks < nks; ++ks) {
      prod_tile(smem, b, rg, m, (ks_first + ks) * 64, (n_first + nt) * TILE_N);
      prod_tile(smem, b, rg, m, (ks_first + ks) * 64, opad + (n_first + nt) * TILE_N);
    }
}
// D[dcol0 + 128 nt ...] (+)= A . W^T for `ntiles` x `nks`; A = hi plane at a_base (k-subtiles ACT_SUB apart), lo plane at + a_lo
__device__ __forceinline__ void mma_gemm(uint8_t* smem, Barriers* b, Ring& rg, uint32_t tmem_base, uint32_t a_base, uint32_t a_lo, int nks,
                                         int dcol0, int ntiles, uint32_t idesc, bool accumulate) {
  for (int nt = 0; nt < ntiles; ++nt) {
    const uint32_t d = tmem_base + dcol0 + nt * TILE_N;
    for (int ks = 0; ks < nks; ++ks) {
      const uint64_t dah = tc::make_smem_desc_sw128(a_base + ks * ACT_SUB, 16, 1024);
      const uint64_t dal = tc::make_smem_desc_sw128(a_base + a_lo + ks * ACT_SUB, 16, 1024);
      {
        const int s = rg.it % NSLOT;
        tc::mbar_wait(&b->full[s], (rg.it / NSLOT) & 1);
        tc::tc_fence_after();
        const uint64_t dbh = tc::make_smem_desc_sw128(tc::smem_u32(smem + OFF_RING + s * W_TILE), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc::umma_bf16(d, dah + 2 * k, dbh + 2 * k, idesc, (accumulate || ks != 0 || k != 0) ? 1u : 0u);
          tc::umma_bf16(d, dal + 2 * k, dbh + 2 * k, idesc, 1);
        }
        tc::umma_commit(&b->empty[s]);
        ++rg.it;
      }
      {
        const int s = rg.it % NSLOT;
        tc::mbar_wait(&b->full[s], (rg.it / NSLOT) & 1);
        tc::tc_fence_after();
        const uint64_t dbl = tc::make_smem_desc_sw128(tc::smem_u32(smem + OFF_RING + s * W_TILE), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc::umma_bf16(d, dah + 2 * k, dbl + 2 * k, idesc, 1);
        tc::umma_commit(&b->empty[s]);
        ++rg.it;
      }
    }
  }
}

// 32 fp32 values of row `row`, columns [32 j, 32 j + 32) -> fp16 hi/lo operand planes (K-major, 128-byte swizzle)
__device__ __forceinline__ void store_operand32(uint8_t* base, int lo_off, int row, int j, const float* v) {
  uint8_t* sub = base + (j >> 1) * ACT_SUB + row * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __half h0, l0, h1, l1;
      split_bf16(v[8 * q + 2 * e], h0, l0); split_bf16(v[8 * q + 2 * e + 1], h1, l1);
      __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
      hi[e] = *reinterpret_cast<uint32_t*>(&hh); lo[e] = *reinterpret_cast<uint32_t*>(&ll);
    }
    const int phys = ((((j & 1) * 4 + q) ^ (row & 7))) * 16;
    *reinterpret_cast<uint4*>(sub + phys) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(sub + lo_off + phys) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
__device__ __forceinline__ void ldg32(const float* __restrict__ p, float* v) {     // 32 consecutive floats (16-byte aligned)
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
  }
}
// coherent variant (plain ld.global): rows this kernel wrote earlier itself must not go through the non-coherent path
__device__ __forceinline__ void ldc32(const float* p, float* v) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 a;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p + 4 * q) : "memory");
    v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
  }
}
__device__ __forceinline__ void stg32(float* p, const float* v) {
#pragma unroll
  for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
// fp32 rows [N][256] of one frame -> ACT operand planes; cooperative over the 128 epilogue threads (coalesced rows)
__device__ __forceinline__ void load_rows_to_act(uint8_t* smem, const float* __restrict__ src, int N, int et /*0..127*/) {
  const int w = et >> 5, lane = et & 31;
  for (int row = w; row < N; row += 4) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int col = half * 128 + lane * 4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + (long)row * C + col));
      __half h0, l0, h1, l1, h2, l2, h3, l3;
      split_bf16(a.x, h0, l0); split_bf16(a.y, h1, l1); split_bf16(a.z, h2, l2); split_bf16(a.w, h3, l3);
      __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3), la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
      const int ks = col >> 6, cidx = (col & 63) >> 3;
      uint8_t* dst = smem + OFF_ACT + ks * ACT_SUB + row * 128 + ((cidx ^ (row & 7)) * 16) + (col & 7) * 2;
      *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&ha), *reinterpret_cast<uint32_t*>(&hb));
      *reinterpret_cast<uint2*>(dst + ACT_PLANE) = make_uint2(*reinterpret_cast<uint32_t*>(&la), *reinterpret_cast<uint32_t*>(&lb));
    }
  }
}

// Epilogue-side view of TMEM for thread r (lane quadrant of its warp)
struct Tm {
  uint32_t base;                                // tmem_base + lane quadrant
  __device__ __forceinline__ void ld(int col, float* v) const { tc::tmem_ld32(base + col, v); tc::tmem_ld_wait(); }
  // raw GEMM output (weights carry WSCALE)
  __device__ __forceinline__ void ldw(int col, float* v) const {
    tc::tmem_ld32(base + col, v); tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] *= WSCALE_INV;
  }
  __device__ __forceinline__ void st(int col, const float* v) const { tc::tmem_st32(base + col, v); }
};
// variance pass + normalisation constants of 256 TMEM columns holding the final pre-norm values (sum already known)
__device__ __forceinline__ float ln_rstd(const Tm& tm, int col0, float mean) {
  float q = 0.f;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    float v[32];
    tm.ld(col0 + 32 * j, v);
#pragma unroll
    for (int c = 0; c < 32; ++c) { const float d = v[c] - mean; q = fmaf(d, d, q); }
  }
  return rsqrtf(q * (1.f / C) + LN_EPS);
}

struct PreParams {
  int N;
  const float *mo, *slots;                      // [T][N][256] MHA output (heads concatenated), slots entering the stage
  const float *out_b, *n1_w, *n1_b, *q_b, *nq_w, *nq_b, *nk_w, *nk_b, *bk_c;
  float *p, *G, *g0, *g1;                       // [T][N][256], [T][N][256], [T][N], [T][N]
  __half* gplanes;                              // [T][2][104][256] hi / lo planes of G (rows >= N zero)
};

__global__ void __launch_bounds__(THREADS, 1)
slot_pre_kernel(const __grid_constant__ CUtensorMap m_out, const __grid_constant__ CUtensorMap m_q, const __grid_constant__ CUtensorMap m_wk,
                const PreParams P) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  Barriers* b = reinterpret_cast<Barriers*>(smem + OFF_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = blockIdx.x, N = P.N;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&m_out); tc::tma_prefetch_desc(&m_q); tc::tma_prefetch_desc(&m_wk);
    for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&b->full[i], 1); tc::mbar_init(&b->empty[i], 1); }
    tc::mbar_init(&b->dfull, 1); tc::mbar_init(&b->aready, 128);
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc(&b->tmem_ptr, 256); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = b->tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      Ring rg;
      prod_gemm(smem, b, rg, &m_out, C, 0, 2, 0, 4);
      prod_gemm(smem, b, rg, &m_q, C, 0, 2, 0, 4);
      prod_gemm(smem, b, rg, &m_wk, C, 0, 2, 0, 4);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      Ring rg;
      const uint32_t act = tc::smem_u32(smem + OFF_ACT);
      for (int g = 0; g < 3; ++g) {
        tc::mbar_wait(&b->aready, g & 1);
        tc::tc_fence_after();
        mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D, 2, IDESC128, false);
        tc::umma_commit(&b->dfull);
      }
    }
  } else {
    const int et = threadIdx.x - 64, q = warp & 3, r = q * 32 + lane;
    const bool valid = r < N;
    Tm tm{tmem_base + ((uint32_t)(q * 32) << 16)};
    const long row = ((long)t * N + r) * C;
    // A operand of the first GEMM: the self-attention output rows
    load_rows_to_act(smem, P.mo + (long)t * N * C, N, et);
    tc::fence_proxy_async();
    tc::mbar_arrive(&b->aready);
    // ---- (1) out_proj + residual + norm1 -> p ----
    tc::mbar_wait(&b->dfull, 0);
    tc::tc_fence_after();
    float s = 0.f;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float v[32], bb[32], rr[32];
      tm.ldw(TM_D + 32 * j, v);
      ldg32(P.out_b + 32 * j, bb);
      if (valid) ldg32(P.slots + row + 32 * j, rr);
#pragma unroll
      for (int c = 0; c < 32; ++c) { v[c] = v[c] + bb[c] + (valid ? rr[c] : 0.f); s += v[c]; }
      tm.st(TM_D + 32 * j, v);
    }
    tc::tmem_st_wait();
    float mean = s * (1.f / C), rstd = ln_rstd(tm, TM_D, mean);
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float v[32], w[32], bb[32];
      tm.ld(TM_D + 32 * j, v);
      ldg32(P.n1_w + 32 * j, w); ldg32(P.n1_b + 32 * j, bb);
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = (v[c] - mean) * rstd * w[c] + bb[c];
      if (valid) { stg32(P.p + row + 32 * j, v); store_operand32(smem + OFF_ACT, ACT_PLANE, r, j, v); }
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&b->aready);
    // ---- (2) to_q + norm_q -> q; qt = q * gamma_k; g0 = qt . bk_c; g1 = q . beta_k ----
    tc::mbar_wait(&b->dfull, 1);
    tc::tc_fence_after();
    s = 0.f;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float v[32], bb[32];
      tm.ldw(TM_D + 32 * j, v);
      ldg32(P.q_b + 32 * j, bb);
#pragma unroll
      for (int c = 0; c < 32; ++c) { v[c] += bb[c]; s += v[c]; }
      tm.st(TM_D + 32 * j, v);
    }
    tc::tmem_st_wait();
    mean = s * (1.f / C); rstd = ln_rstd(tm, TM_D, mean);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float v[32], w[32], bb[32], gk[32], bk[32], bc[32];
      tm.ld(TM_D + 32 * j, v);
      ldg32(P.nq_w + 32 * j, w); ldg32(P.nq_b + 32 * j, bb);
      ldg32(P.nk_w + 32 * j, gk); ldg32(P.nk_b + 32 * j, bk); ldg32(P.bk_c + 32 * j, bc);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float qv = (v[c] - mean) * rstd * w[c] + bb[c];
        const float tv = qv * gk[c];
        s0 = fmaf(tv, bc[c], s0);
        s1 = fmaf(qv, bk[c], s1);
        v[c] = tv;
      }
      if (valid) store_operand32(smem + OFF_ACT, ACT_PLANE, r, j, v);
    }
    if (valid) { P.g0[(long)t * N + r] = s0; P.g1[(long)t * N + r] = s1; }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&b->aready);
    // ---- (3) G = qt . Wk_c -> fp32 + fp16 hi/lo planes (the B operand of attn_tc's S product) ----
    tc::mbar_wait(&b->dfull, 0);
    tc::tc_fence_after();
    __half* gp = P.gplanes + ((long)t * 2 * NR + r) * C;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float v[32];
      tm.ldw(TM_D + 32 * j, v);
      if (r < NR) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          __half h0 = __float2half_rn(0.f), l0 = h0, h1 = h0, l1 = h0;
          if (valid) { split_bf16(v[2 * c], h0, l0); split_bf16(v[2 * c + 1], h1, l1); }
          __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
          hi[c] = *reinterpret_cast<uint32_t*>(&hh); lo[c] = *reinterpret_cast<uint32_t*>(&ll);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          reinterpret_cast<uint4*>(gp + 32 * j)[c] = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          reinterpret_cast<uint4*>(gp + (long)NR * C + 32 * j)[c] = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
      }
      if (valid && P.G) stg32(P.G + row + 32 * j, v);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 256); }
}

// ---- post-attention phase -----------------------------------------------------------------------------------------
struct PostParams {
  int N, F, act, phases, ncls;                  // phases: bit 0 = attention epilogue + FFN, bit 1 = towers
  // phase 0
  const float *Z, *a0, *a1, *p;
  const float *nv_w, *nv_b, *bv_c, *no_w, *no_b, *n2_w, *n2_b, *b1, *b2, *n3_w, *n3_b;
  float *p2buf, *f_out;                         // [T][N][256] scratch (post-norm2 rows), FFN block output
  // towers
  const float* f_in;                            // [T][N][256] tower input when phase 0 did not just produce it in shared memory
  const float *tw_ln_w, *tw_ln_b, *c1_nw, *c1_nb, *r1_nw, *r1_nb, *logit_b;
  float *slots_out, *emb_out, *cls_out;
  long emb_fs, cls_fs;                          // frame strides of emb_out / cls_out
};

__device__ __forceinline__ float act_fn(float x, int act) { return act == 1 ? fmaxf(x, 0.f) : gelu_erf(x); }

__global__ void __launch_bounds__(THREADS, 1)
slot_post_kernel(const __grid_constant__ CUtensorMap m_wv, const __grid_constant__ CUtensorMap m_l1, const __grid_constant__ CUtensorMap m_l2,
                 const __grid_constant__ CUtensorMap m_tw, const __grid_constant__ CUtensorMap m_c1, const __grid_constant__ CUtensorMap m_r1,
                 const __grid_constant__ CUtensorMap m_lg, const PostParams P) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  Barriers* b = reinterpret_cast<Barriers*>(smem + OFF_MISC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = blockIdx.x, N = P.N;
  const int NC = P.F / 128;                     // hidden chunks
  const bool ph0 = P.phases & 1, ph1 = (P.phases & 2) != 0;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&m_wv); tc::tma_prefetch_desc(&m_l1); tc::tma_prefetch_desc(&m_l2);
    tc::tma_prefetch_desc(&m_tw); tc::tma_prefetch_desc(&m_c1); tc::tma_prefetch_desc(&m_r1); tc::tma_prefetch_desc(&m_lg);
    for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&b->full[i], 1); tc::mbar_init(&b->empty[i], 1); }
    tc::mbar_init(&b->dfull, 1); tc::mbar_init(&b->aready, 128);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&b->d1full[i], 1); tc::mbar_init(&b->d1free[i], 128); }
    tc::mbar_init(&b->hfull, 128); tc::mbar_init(&b->hfree, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc(&b->tmem_ptr, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = b->tmem_ptr;
  const int F_pad = P.F;                        // lin1 planes [2][F][256] (F is a multiple of 128)

  if (warp == 0) {
    // ===================== TMA producer: the stage's weights in consumption order =====================
    if (lane == 0) {
      Ring rg;
      if (ph0) {
        prod_gemm(smem, b, rg, &m_wv, C, 0, 2, 0, 4);
        prod_gemm(smem, b, rg, &m_l1, F_pad, 0, 1, 0, 4);
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) prod_gemm(smem, b, rg, &m_l1, F_pad, c + 1, 1, 0, 4);
          prod_gemm(smem, b, rg, &m_l2, C, 0, 2, 2 * c, 2);
        }
      }
      if (ph1) {
        prod_gemm(smem, b, rg, &m_tw, 2 * C, 0, 4, 0, 4);
        prod_gemm(smem, b, rg, &m_c1, C, 0, 2, 0, 4);
        prod_gemm(smem, b, rg, &m_lg, TILE_N, 0, 1, 0, 4);
        prod_gemm(smem, b, rg, &m_r1, C, 0, 2, 0, 4);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      Ring rg;
      const uint32_t act = tc::smem_u32(smem + OFF_ACT), hb = tc::smem_u32(smem + OFF_HB);
      uint32_t na = 0;                                            // uses of the aready barrier
      auto wait_act = [&]() { tc::mbar_wait(&b->aready, na & 1); ++na; tc::tc_fence_after(); };
      if (ph0) {
        wait_act();                                               // Z rows are in ACT
        mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D, 2, IDESC128, false);      // Y = Z . Wv_c^T
        tc::umma_commit(&b->dfull);
        wait_act();                                               // p2 rows are in ACT, D is drained
        mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D1, 1, IDESC128, false);     // hidden chunk 0
        tc::umma_commit(&b->d1full[0]);
        for (int c = 0; c < NC; ++c) {
          if (c + 1 < NC) {
            const int bf = (c + 1) & 1;
            if (c + 1 >= 2) { tc::mbar_wait(&b->d1free[bf], (((c + 1) >> 1) - 1) & 1); tc::tc_fence_after(); }
            mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D1 + bf * 128, 1, IDESC128, false);
            tc::umma_commit(&b->d1full[bf]);
          }
          tc::mbar_wait(&b->hfull, c & 1);                        // activated hidden chunk c is in HB
          tc::tc_fence_after();
          mma_gemm(smem, b, rg, tmem_base, hb, HB_PLANE, 2, TM_D, 2, IDESC128, c != 0);      // out += h_c . W2[:, chunk]^T
          tc::umma_commit(&b->hfree);
        }
        tc::umma_commit(&b->dfull);
      }
      if (ph1) {
        wait_act();                                               // tower input rows in ACT (D drained)
        mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D, 4, IDESC128, false);      // cls0 | reg0 -> 512 columns
        tc::umma_commit(&b->dfull);
        wait_act();                                               // c1 in ACT, columns [0,256) drained
        mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D, 2, IDESC128, false);      // cls1
        tc::umma_commit(&b->dfull);
        wait_act();                                               // c2 in ACT
        mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D, 1, IDESC32, false);       // class logits (32 padded columns)
        tc::umma_commit(&b->dfull);
        wait_act();                                               // e1 in ACT
        mma_gemm(smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D, 2, IDESC128, false);      // reg1
        tc::umma_commit(&b->dfull);
      }
    }
  } else {
    // ===================== epilogue: thread r owns slot row r =====================
    const int et = threadIdx.x - 64, q = warp & 3, r = q * 32 + lane;
    const bool valid = r < N;
    Tm tm{tmem_base + ((uint32_t)(q * 32) << 16)};
    const long row = ((long)t * N + r) * C;
    uint32_t nd = 0;                                              // uses of the dfull barrier
    auto wait_d = [&]() { tc::mbar_wait(&b->dfull, nd & 1); ++nd; tc::tc_fence_after(); };
    auto publish_act = [&]() { tc::tc_fence_before(); tc::fence_proxy_async(); tc::mbar_arrive(&b->aready); };
    // y = act(LN(x)) over 256 TMEM columns at col0 whose pre-norm sum is `s`; emit(j, v) receives 32 normalised values
    auto ln_emit = [&](int col0, float s, const float* gw, const float* gb, bool relu, auto&& emit) {
      const float mean = s * (1.f / C), rstd = ln_rstd(tm, col0, mean);
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        float v[32], w[32], bb[32];
        tm.ld(col0 + 32 * j, v);
        ldg32(gw + 32 * j, w); ldg32(gb + 32 * j, bb);
#pragma unroll
        for (int c = 0; c < 32; ++c) { v[c] = (v[c] - mean) * rstd * w[c] + bb[c]; if (relu) v[c] = fmaxf(v[c], 0.f); }
        emit(j, v);
      }
    };
    if (ph0) {
      load_rows_to_act(smem, P.Z + (long)t * N * C, N, et);
      publish_act();
      // ---- value projection of the pixel-reduced slots, norm_v / norm1 / ReLU, residual, norm2 (:456-459, 374-376) ----
      wait_d();
      const float a0r = valid ? P.a0[(long)t * N + r] : 0.f, a1r = valid ? P.a1[(long)t * N + r] : 0.f;
      float s = 0.f;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        float v[32], gv[32], bv[32], bc[32];
        tm.ldw(TM_D + 32 * j, v);
        ldg32(P.nv_w + 32 * j, gv); ldg32(P.nv_b + 32 * j, bv); ldg32(P.bv_c + 32 * j, bc);
#pragma unroll
        for (int c = 0; c < 32; ++c) { v[c] = gv[c] * fmaf(bc[c], a1r, v[c]) + bv[c] * a0r; s += v[c]; }
        tm.st(TM_D + 32 * j, v);
      }
      tc::tmem_st_wait();
      float s2 = 0.f;
      ln_emit(TM_D, s, P.no_w, P.no_b, true, [&](int j, float* v) {
        float pp[32];
        if (valid) ldg32(P.p + row + 32 * j, pp);
#pragma unroll
        for (int c = 0; c < 32; ++c) { v[c] = (valid ? pp[c] : 0.f) + v[c]; s2 += v[c]; }
        tm.st(TM_D + 32 * j, v);
      });
      tc::tmem_st_wait();
      ln_emit(TM_D, s2, P.n2_w, P.n2_b, false, [&](int j, float* v) {
        if (valid) { stg32(P.p2buf + row + 32 * j, v); store_operand32(smem + OFF_ACT, ACT_PLANE, r, j, v); }
      });
      publish_act();
      // ---- FFN: hidden chunks of 128 (:379-382) ----
      for (int c = 0; c < NC; ++c) {
        const int bf = c & 1;
        tc::mbar_wait(&b->d1full[bf], (c >> 1) & 1);
        tc::tc_fence_after();
        float h[4][32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float bb[32];
          tc::tmem_ld32(tm.base + TM_D1 + bf * 128 + 32 * j, h[j]);
          ldg32(P.b1 + c * 128 + 32 * j, bb);
          tc::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) h[j][e] = act_fn(fmaf(h[j][e], WSCALE_INV, bb[e]), P.act);
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&b->d1free[bf]);
        if (c > 0) tc::mbar_wait(&b->hfree, (c - 1) & 1);         // the MMAs of chunk c-1 have finished reading HB
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) store_operand32(smem + OFF_HB, HB_PLANE, r, j, h[j]);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&b->hfull);
      }
      // ---- linear2 bias + residual + norm3 -> f ----
      wait_d();
      s = 0.f;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        float v[32], bb[32], pp[32];
        tm.ldw(TM_D + 32 * j, v);
        ldg32(P.b2 + 32 * j, bb);
        if (valid) ldc32(P.p2buf + row + 32 * j, pp);
#pragma unroll
        for (int c = 0; c < 32; ++c) { v[c] = v[c] + bb[c] + (valid ? pp[c] : 0.f); s += v[c]; }
        tm.st(TM_D + 32 * j, v);
      }
      tc::tmem_st_wait();
      ln_emit(TM_D, s, P.n3_w, P.n3_b, false, [&](int j, float* v) {
        if (valid) { stg32(P.f_out + row + 32 * j, v); if (ph1) store_operand32(smem + OFF_ACT, ACT_PLANE, r, j, v); }
      });
      if (ph1) publish_act();
    } else if (ph1) {
      load_rows_to_act(smem, P.f_in + (long)t * N * C, N, et);
      publish_act();
    }
    if (ph1) {
      // ---- towers (:390-400): first layers cls0 | reg0 share the input; 512 accumulator columns ----
      wait_d();
      auto col_sum = [&](int col0) {                              // unscale the raw GEMM output in place, return the row sum
        float s = 0.f;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          float v[32];
          tm.ldw(col0 + 32 * j, v);
#pragma unroll
          for (int c = 0; c < 32; ++c) s += v[c];
          tm.st(col0 + 32 * j, v);
        }
        tc::tmem_st_wait();
        return s;
      };
      float s = col_sum(TM_D);
      ln_emit(TM_D, s, P.tw_ln_w, P.tw_ln_b, true, [&](int j, float* v) { if (valid) store_operand32(smem + OFF_ACT, ACT_PLANE, r, j, v); });
      publish_act();                                              // c1 -> cls1 GEMM (writes columns [0,256); the reg half stays)
      wait_d();
      s = col_sum(TM_D);
      ln_emit(TM_D, s, P.c1_nw, P.c1_nb, true, [&](int j, float* v) { if (valid) store_operand32(smem + OFF_ACT, ACT_PLANE, r, j, v); });
      publish_act();                                              // c2 -> class logits
      wait_d();
      {
        float v[32];
        tm.ldw(TM_D, v);
        if (valid) {
          float* dst = P.cls_out + (long)t * P.cls_fs + (long)r * P.ncls;
          for (int c = 0; c < P.ncls; ++c) dst[c] = v[c] + __ldg(P.logit_b + c);
        }
      }
      s = col_sum(TM_D + 256);                                    // reg half of the first tower layer
      ln_emit(TM_D + 256, s, P.tw_ln_w + C, P.tw_ln_b + C, true, [&](int j, float* v) { if (valid) store_operand32(smem + OFF_ACT, ACT_PLANE, r, j, v); });
      publish_act();                                              // e1 -> reg1 GEMM
      wait_d();
      s = col_sum(TM_D);
      ln_emit(TM_D, s, P.r1_nw, P.r1_nb, true, [&](int j, float* v) {
        if (valid) { stg32(P.slots_out + row + 32 * j, v); stg32(P.emb_out + (long)t * P.emb_fs + (long)r * C + 32 * j, v); }
      });
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// ---- host side ----------------------------------------------------------------------------------------------------
struct SlotTcWeights {                          // fp16 hi/lo planes [2][Opad][K] per linear layer of one stage
  __half *out_proj, *to_q, *wkT, *wv, *lin1, *lin2, *tw, *cls1, *reg1, *logit;
};
inline void slot_tc_layout(Arena& a, const slotvps_head_desc* d, SlotTcWeights* w) {
  const size_t cc = (size_t)2 * C * C;
  w->out_proj = a.take<__half>(cc); w->to_q = a.take<__half>(cc); w->wkT = a.take<__half>(cc); w->wv = a.take<__half>(cc);
  w->lin1 = a.take<__half>((size_t)2 * d->dim_feedforward * C); w->lin2 = a.take<__half>((size_t)2 * C * d->dim_feedforward);
  w->tw = a.take<__half>(2 * cc); w->cls1 = a.take<__half>(cc); w->reg1 = a.take<__half>(cc);
  w->logit = a.take<__half>((size_t)2 * slot::TILE_N * C);
}
inline bool slot_tc_supported(const slotvps_head_desc* d) {
  return d->kernel_path == 0 && d->n_slots <= slot::NR && d->dim_feedforward % 128 == 0 && d->num_classes <= 32;
}
inline int slot_planes(const float* W, int O, int K, __half* out, cudaStream_t s) {
  const int Opad = ceil_div(O, slot::TILE_N) * slot::TILE_N;
  slot::linear_planes_kernel<<<(unsigned)(((long)Opad * K + 255) / 256), 256, 0, s>>>(W, O, K, Opad, out);
  SV_CHECK_LAUNCH("linear_planes");
  return SLOTVPS_OK;
}
// tensor map over [2 * Opad][K] fp16 planes, box [128][64]
inline int slot_wmap(CUtensorMap* m, const __half* planes, int O, int K) {
  const int Opad = ceil_div(O, slot::TILE_N) * slot::TILE_N;
  return tc::make_tmap_h16_sw128(m, planes, (uint64_t)2 * Opad, (uint64_t)K, slot::TILE_N);
}

}  // namespace slot
}  // namespace slotvps
