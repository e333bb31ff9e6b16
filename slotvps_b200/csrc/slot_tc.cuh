// Slot-update kernels on the tensor pipe (north_star kernel 2), shared pieces: the slot side of one retriever stage
// (dynamic_mask_head.py:342-400) keeps the N <= 104 slots of a frame on the 128 TMEM lanes of a CTA -- activations live
// in shared memory as fp16 hi/lo MMA operands and in TMEM as fp32 accumulators, weights stream through a TMA ring, every
// LayerNorm / GELU / residual runs on slot rows straight out of TMEM.  GEMM scheme: fp16 hi/lo operands, 3 products,
// fp32 accumulation.
//
//   slot_cl.cuh       the row-wise layers on a cluster of four CTAs per frame (64 output columns each)
//   slot_ffn_kernel   the FFN 256 -> F -> 256 over (frame, 128-wide hidden chunk) CTAs   (this file)
//   slot_norm3_kernel chunk-ordered reduction of the FFN partials + residual + norm3      (this file)
//
// Operand scaling.  fp16 has 5 exponent bits: the lo plane of a value below ~0.1 is subnormal (absolute step 6e-8), i.e.
// the hi/lo pair carries ~1e-6 relative instead of 2^-22 -- and for sums of random-sign terms the relative error of the sum
// equals the per-term one.  Weight planes are therefore stored times 2^8, LayerNorm outputs are written times 2^4, and rows
// loaded from global memory (attention outputs, pixel-reduced slots, GELU outputs: anything from 1e-3 to 1e4) get a per-row
// power of two that puts the row maximum in [512, 1024); the accumulator read-back undoes the scales exactly.  Measured:
// per-stage teacher-forced error 3.0e-6 -> 1.4e-6 (the fp32 path: 1.36e-6).
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
constexpr int MISC_BYTES = 1024;               // barriers, tmem pointer
constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
constexpr int EPI_WARPS = 16, EPI_THREADS = 512;
constexpr int THREADS = 64 + EPI_THREADS;      // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr uint32_t IDESC128 = tc::make_idesc_f16(128, 128, 0, 0);
constexpr uint32_t IDESC32 = tc::make_idesc_f16(128, 32, 0, 0);
constexpr int TM_D = 0;                        // main accumulator, 256 columns (towers: 512)
constexpr int TM_D1 = 256;                     // two 128-column FFN hidden accumulators
// Weight planes are stored scaled by 2^8: fp16 has 5 exponent bits, so the lo plane of a weight of ~0.05 would be
// subnormal (absolute step 6e-8, ~1e-6 relative instead of 2.4e-7 -- and for sums of random-sign terms the relative
// error of the sum equals the per-term one).  Every accumulator read is multiplied by 2^-8 (exact).
constexpr float WSCALE = 256.f, WSCALE_INV = 1.f / 256.f;
constexpr float LSCALE = 16.f, LSCALE_INV = 1.f / 16.f;      // LayerNorm outputs (|x| <= 16 |gamma| + |beta|) as operands
constexpr int OFF_RSC = OFF_MISC + 512;        // [128] floats: inverse row scales of the rows loaded by load_rows_to_act
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(OFF_HB % 1024 == 0 && OFF_RING % 1024 == 0 && ACT_SUB % 1024 == 0, "swizzle atoms");

// A frame's N slots are processed as G = ceil(N / 104) groups of per = ceil(N / G) consecutive rows (the last one ragged):
// every layer of the slot update is row-wise, so a group is simply a "virtual frame" vf = t * G + g.
struct RowGroup { int t, row0, n; long rb; };   // frame, first slot, slots in the group, first global row (t * N + row0)
__host__ __device__ __forceinline__ RowGroup row_group(int vf, int N, int G) {
  RowGroup r;
  const int per = (N + G - 1) / G, g = vf % G;
  r.t = vf / G; r.row0 = g * per; r.n = min(per, N - r.row0); r.rb = (long)r.t * N + r.row0;
  return r;
}
inline int slot_groups(int N) { return (N + NR - 1) / NR; }

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
// all lanes of the MMA warp run this (warp-uniform waits and descriptors); `el` = tc::elect_one() guards the tcgen05 instructions
__device__ __forceinline__ void mma_gemm(bool el, uint8_t* smem, Barriers* b, Ring& rg, uint32_t tmem_base, uint32_t a_base, uint32_t a_lo, int nks,
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
        if (el) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc::umma_bf16(d, dah + 2 * k, dbh + 2 * k, idesc, (accumulate || ks != 0 || k != 0) ? 1u : 0u);
            tc::umma_bf16(d, dal + 2 * k, dbh + 2 * k, idesc, 1);
          }
          tc::umma_commit(&b->empty[s]);
        }
        ++rg.it;
      }
      {
        const int s = rg.it % NSLOT;
        tc::mbar_wait(&b->full[s], (rg.it / NSLOT) & 1);
        tc::tc_fence_after();
        const uint64_t dbl = tc::make_smem_desc_sw128(tc::smem_u32(smem + OFF_RING + s * W_TILE), 16, 1024);
        if (el) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_bf16(d, dah + 2 * k, dbl + 2 * k, idesc, 1);
          tc::umma_commit(&b->empty[s]);
        }
        ++rg.it;
      }
    }
  }
}

// L2 prefetch of the same tiles (no shared-memory destination), issued ahead of the 4-slot ring
__device__ __forceinline__ void prefetch_gemm(const CUtensorMap* m, int opad, int n_first, int ntiles, int ks_first, int nks) {
  for (int nt = 0; nt < ntiles; ++nt)
    for (int ks = 0; ks < nks; ++ks) {
      tc::tma_prefetch_2d(m, (ks_first + ks) * 64, (n_first + nt) * TILE_N);
      tc::tma_prefetch_2d(m, (ks_first + ks) * 64, opad + (n_first + nt) * TILE_N);
    }
}

// 16 fp32 values of row `row`, columns [16 u, 16 u + 16) -> fp16 hi/lo operand planes (K-major, 128-byte swizzle)
__device__ __forceinline__ void store_operand16(uint8_t* base, int lo_off, int row, int u, const float* v, float sc) {
  uint8_t* sub = base + (u >> 2) * ACT_SUB + row * 128;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(v[8 * q + 2 * e] * sc, v[8 * q + 2 * e + 1] * sc, hi[e], lo[e]);
    const int phys = ((((u & 3) * 2 + q) ^ (row & 7))) * 16;
    *reinterpret_cast<uint4*>(sub + phys) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(sub + lo_off + phys) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
// power-of-two scale that puts a row maximum m >= 0 into [512, 1024) (capped at 2^40 for tiny rows); inv = its inverse
__device__ __forceinline__ float row_scale(float m, float& inv) {
  const int ex = max((__float_as_int(m) >> 23) & 0xFF, 96);
  inv = __int_as_float((ex - 9) << 23);
  return __int_as_float((263 - ex) << 23);
}
// fp32 rows [N][256] of one frame -> ACT operand planes, each row times its own power of two (inverse -> rsc[row]);
// cooperative over the epilogue warps (coalesced row reads, a warp owns a row at a time)
__device__ __forceinline__ void load_rows_to_act(uint8_t* smem, const float* __restrict__ src, int N, int we, int lane, float* rsc) {
  // four rows per pass: eight 16-byte loads in flight per lane before the first cross-lane maximum
#pragma unroll 1
  for (int row0 = we; row0 < N; row0 += 4 * EPI_WARPS) {
    float4 a[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = min(row0 + i * EPI_WARPS, N - 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) a[i][half] = __ldg(reinterpret_cast<const float4*>(src + (long)row * C + half * 128 + lane * 4));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + i * EPI_WARPS;
      if (row >= N) break;
      float m = fmaxf(fmaxf(fmaxf(fabsf(a[i][0].x), fabsf(a[i][0].y)), fmaxf(fabsf(a[i][0].z), fabsf(a[i][0].w))),
                      fmaxf(fmaxf(fabsf(a[i][1].x), fabsf(a[i][1].y)), fmaxf(fabsf(a[i][1].z), fabsf(a[i][1].w))));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float inv;
      const float sc = row_scale(m, inv);
      if (lane == 0) rsc[row] = inv;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = half * 128 + lane * 4;
        uint32_t h0, l0, h1, l1;
        split2(a[i][half].x * sc, a[i][half].y * sc, h0, l0); split2(a[i][half].z * sc, a[i][half].w * sc, h1, l1);
        const int ks = col >> 6, cidx = (col & 63) >> 3;
        uint8_t* dst = smem + OFF_ACT + ks * ACT_SUB + row * 128 + ((cidx ^ (row & 7)) * 16) + (col & 7) * 2;
        *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(dst + ACT_PLANE) = make_uint2(l0, l1);
      }
    }
  }
}
__device__ __forceinline__ void ldg16(const float* __restrict__ p, float* v) {     // 16 consecutive floats (16-byte aligned)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
  }
}

// Epilogue context: 16 warps; warp e serves TMEM lane quadrant (e & 3) -- slot rows 32 (e & 3) + lane -- and column
// quarter (e >> 2) of the accumulator, so a slot row is shared by four threads.  The epilogues are instruction-bound,
// hence the many warps.  Global row I/O goes through a per-warp staging tile so that every request is coalesced (a
// thread-per-row access pattern costs ~100 cycles per instruction: 21 K cycles per epilogue were measured with it).
struct Epi {
  int r, qt, lane, q, N;                        // row, column quarter, lane, lane quadrant
  bool valid;
  uint32_t tbase;                               // tmem_base + lane quadrant
  float* stg;                                   // [32][20] staging tile of this warp
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync 1, 512;" ::: "memory"); }
  // accumulator columns of unit u (16 columns) times `inv` (the product of the inverse operand scales)
  __device__ __forceinline__ void ld_raw(int col0, int u, float* v, float inv) const {
    tc::tmem_ld16(tbase + col0 + 16 * u, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] *= inv;
  }
  // v[c] = frame[(32 q + lane) * 256 + 16 u + c] (zeros for rows >= N)
  __device__ __forceinline__ void read_rows(const float* frame, int u, float* v, bool coherent = false) const {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2), gr = q * 32 + row, c4 = (lane & 3) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < N) {
        const float* src = frame + (long)gr * C + 16 * u + c4;
        if (coherent) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src) : "memory");
        else a = __ldg(reinterpret_cast<const float4*>(src));
      }
      *reinterpret_cast<float4*>(stg + row * 20 + c4) = a;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(stg + lane * 20 + 4 * c);
      v[4 * c] = a.x; v[4 * c + 1] = a.y; v[4 * c + 2] = a.z; v[4 * c + 3] = a.w;
    }
    __syncwarp();
  }
  __device__ __forceinline__ void write_rows(float* frame, int u, const float* v, float* frame2 = nullptr, int ld = C) const {
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(stg + lane * 20 + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2), gr = q * 32 + row, c4 = (lane & 3) * 4;
      if (gr < N) {
        const float4 a = *reinterpret_cast<const float4*>(stg + row * 20 + c4);
        *reinterpret_cast<float4*>(frame + (long)gr * ld + 16 * u + c4) = a;
        if (frame2) *reinterpret_cast<float4*>(frame2 + (long)gr * ld + 16 * u + c4) = a;
      }
    }
    __syncwarp();
  }
};
__device__ __forceinline__ Epi make_epi(uint8_t* smem, uint32_t tmem_base, int N) {
  Epi e;
  const int we = (threadIdx.x >> 5) - 2;
  e.lane = threadIdx.x & 31; e.q = (threadIdx.x >> 5) & 3; e.qt = we >> 2; e.r = e.q * 32 + e.lane; e.N = N;
  e.valid = e.r < N;
  e.tbase = tmem_base + ((uint32_t)(e.q * 32) << 16);
  e.stg = reinterpret_cast<float*>(smem + OFF_HB) + we * 640;      // 2560 B per warp: [32][20] floats (HB is idle when rows are written)
  return e;
}

struct PreParams {
  int N, G, dbg;                                // slots per frame, row groups per frame (clusters per frame)
  const float *mo, *slots;                      // [T][N][256] MHA output (heads concatenated), slots entering the stage
  const float *out_b, *n1_w, *n1_b, *q_b, *nq_w, *nq_b, *nk_w, *nk_b, *bk_c;
  float *p, *Gout, *g0, *g1;                    // [T][N][256], [T][N][256], [T][N], [T][N]
  __half* gplanes;                              // [T][2][104][256] hi / lo planes of G (rows >= N zero)
  uint8_t* opx;                                 // [T][106496] operand image exchanged between the CTAs of a frame's cluster
};


// ---- post-attention phase -----------------------------------------------------------------------------------------
struct PostParams {
  int N, G, ncls;
  const float *Z, *a0, *a1, *p;                 // pixel-reduced slots [T][N][256], softmax mass / bias terms [T][N], residual rows
  const float *nv_w, *nv_b, *bv_c, *no_w, *no_b, *n2_w, *n2_b;
  float* p2buf;                                 // [T][N][256] post-norm2 rows (FFN input and residual)
  // towers
  const float* f_in;                            // [T][N][256] tower input
  const float *tw_ln_w, *tw_ln_b, *c1_nw, *c1_nb, *r1_nw, *r1_nb, *logit_b;
  float *slots_out, *emb_out, *cls_out;
  long emb_fs, cls_fs;                          // frame strides of emb_out / cls_out
  uint8_t* opx;                                 // [T][106496] operand image exchanged between the CTAs of a frame's cluster
};

// Exact (erf) GELU through the complementary error function: gelu(x) = x/2 * erfc(-x/sqrt(2)), with erfc(z >= 0) from the
// Chebyshev fit t * exp(-z^2 + P9(t)), t = 1 / (1 + z/2) (fractional error < 1.2e-7 everywhere; W. H. Press et al.).  Against
// fp64 this is 6.1e-7 absolute / 1.7e-6 relative -- tighter than the fp32 form 0.5 x (1 + erf(x / sqrt 2)) the reference
// evaluates (6.7e-7 / 5.4e-5, cancellation for x < 0) -- at ~1/3 of erff's instructions: the 2048-wide hidden activation is
// the largest single item of the slot update (measured 100 K of 176 K FFN cycles per stage with erff).
__device__ __forceinline__ float gelu_erfc(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.5f, z, 1.f)));
  float p = 0.17087277f;
  p = fmaf(p, t, -0.82215223f); p = fmaf(p, t, 1.48851587f); p = fmaf(p, t, -1.13520398f); p = fmaf(p, t, 0.27886807f);
  p = fmaf(p, t, -0.18628806f); p = fmaf(p, t, 0.09678418f); p = fmaf(p, t, 0.37409196f); p = fmaf(p, t, 1.00002368f);
  p = fmaf(p, t, -1.26551223f);
  const float e = t * __expf(fmaf(-z, z, p));             // erfc(|x| / sqrt 2)
  return 0.5f * x * (x >= 0.f ? 2.f - e : e);
}
__device__ __forceinline__ float act_fn(float x, int act) { return act == 1 ? fmaxf(x, 0.f) : gelu_erfc(x); }


// ---- FFN of one stage spread over the GPU (:379-385) ---------------------------------------------------------------------
// One CTA per (frame, 128-wide hidden chunk): the 2 x 3-product GEMMs of the FFN are 100 K tensor-pipe cycles on a single SM
// (M = 128 x 256 x 2048 x 2 x 3) and the activation another ~60 K issue cycles, so the frame-resident kernel spent 3/4 of a
// stage here.  Every CTA rebuilds the fp16 operand planes of the frame's post-norm2 rows (106 KB from L2), runs
// lin1[chunk] -> activation -> lin2[:, chunk] and writes its [N][256] partial; slot_norm3_kernel adds the partials in chunk
// order (deterministic), the bias and the residual and applies norm3.
struct FfnParams {
  int N, G, act;
  const float *p2, *b1;                         // [T][N][256] post-norm2 rows, [F] lin1 bias
  float* part;                                  // [F/128][T][N][256] partial lin2 outputs
  long part_stride;                             // T * N * 256
};

__global__ void __launch_bounds__(THREADS, 1)
slot_ffn_kernel(const __grid_constant__ CUtensorMap m_l1, const __grid_constant__ CUtensorMap m_l2, const FfnParams P, const int F_pad) {
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  Barriers* b = reinterpret_cast<Barriers*>(smem + OFF_MISC);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, c = blockIdx.y;
  const RowGroup rg = row_group(blockIdx.x, P.N, P.G);
  const int N = rg.n;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&m_l1); tc::tma_prefetch_desc(&m_l2);
    for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&b->full[i], 1); tc::mbar_init(&b->empty[i], 1); }
    tc::mbar_init(&b->dfull, 1); tc::mbar_init(&b->aready, EPI_THREADS);
    tc::mbar_init(&b->d1full[0], 1); tc::mbar_init(&b->hfull, EPI_THREADS);
    tc::fence_barrier_init();
  }
  if (warp == 1) { tc::tmem_alloc(&b->tmem_ptr, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, b->tmem_ptr, 0);
  if (warp == 0) {
    if (lane == 0) {
      Ring rg;
      prod_gemm(smem, b, rg, &m_l1, F_pad, c, 1, 0, 4);
      prod_gemm(smem, b, rg, &m_l2, C, 0, 2, 2 * c, 2);
    }
  } else if (warp == 1) {
    {                                                               // all lanes: warp-uniform issue loop (see tc::elect_one)
      const bool el = tc::elect_one();
      Ring rg;
      const uint32_t act = tc::smem_u32(smem + OFF_ACT), hb = tc::smem_u32(smem + OFF_HB);
      tc::mbar_wait(&b->aready, 0); tc::tc_fence_after();
      mma_gemm(el, smem, b, rg, tmem_base, act, ACT_PLANE, 4, TM_D1, 1, IDESC128, false);   // hidden chunk
      if (el) tc::umma_commit(&b->d1full[0]);
      tc::mbar_wait(&b->hfull, 0); tc::tc_fence_after();
      mma_gemm(el, smem, b, rg, tmem_base, hb, HB_PLANE, 2, TM_D, 2, IDESC128, false);      // partial out = h_c . W2[:, chunk]^T
      if (el) tc::umma_commit(&b->dfull);
    }
  } else {
    Epi e = make_epi(smem, tmem_base, N);
    const long fbase = rg.rb * C;
    float* rsc = reinterpret_cast<float*>(smem + OFF_RSC);
    load_rows_to_act(smem, P.p2 + fbase, N, warp - 2, lane, rsc);
    tc::tc_fence_before(); tc::fence_proxy_async(); tc::mbar_arrive(&b->aready);
    float bb[2][16];
#pragma unroll
    for (int uu = 0; uu < 2; ++uu) ldg16(P.b1 + c * 128 + 16 * (e.qt * 2 + uu), bb[uu]);
    e.sync();                                                      // rsc is complete
    const float inv1 = WSCALE_INV * rsc[e.r];
    tc::mbar_wait(&b->d1full[0], 0);
    tc::tc_fence_after();
    float h[2][16];
#pragma unroll
    for (int uu = 0; uu < 2; ++uu) tc::tmem_ld16(e.tbase + TM_D1 + 16 * (e.qt * 2 + uu), h[uu]);
    tc::tmem_ld_wait();
    float m = 0.f;
#pragma unroll
    for (int uu = 0; uu < 2; ++uu)
#pragma unroll
      for (int k = 0; k < 16; ++k) { h[uu][k] = act_fn(fmaf(h[uu][k], inv1, bb[uu][k]), P.act); m = fmaxf(m, fabsf(h[uu][k])); }
    // row maximum of the activated chunk over the four threads of the row (lin1 has retired: ACT is free) -> operand scale
    float* hm = reinterpret_cast<float*>(smem + OFF_ACT);
    hm[e.qt * 128 + e.r] = m;
    e.sync();
    m = fmaxf(fmaxf(hm[e.r], hm[128 + e.r]), fmaxf(hm[256 + e.r], hm[384 + e.r]));
    float hinv;
    const float hsc = row_scale(m, hinv);
    if (e.valid) {
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) store_operand16(smem + OFF_HB, HB_PLANE, e.r, e.qt * 2 + uu, h[uu], hsc);
    }
    tc::tc_fence_before(); tc::fence_proxy_async(); tc::mbar_arrive(&b->hfull);
    tc::mbar_wait(&b->dfull, 0);                                   // lin2 retired: HB is free for the staging tiles
    tc::tc_fence_after();
    float* dst = P.part + (long)c * P.part_stride + fbase;
    const float inv2 = WSCALE_INV * hinv;
#pragma unroll 1
    for (int uu = 0; uu < 4; ++uu) {
      float v[16];
      e.ld_raw(TM_D, e.qt * 4 + uu, v, inv2);
      e.write_rows(dst, e.qt * 4 + uu, v);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

// f = [post +] norm3(p2 + b2 + sum_c part[c]) (:382-385; Video Retriever :317 with post = X); one warp per slot row,
// two-pass LayerNorm in registers
__global__ void __launch_bounds__(256) slot_norm3_kernel(const float* __restrict__ part, long part_stride, int nparts, const float* __restrict__ p2,
                                                         const float* __restrict__ b2, const float* __restrict__ gw, const float* __restrict__ gb,
                                                         const float* __restrict__ post, float* __restrict__ f_out, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int col = hh * 128 + lane * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int cc = 0; cc < nparts; ++cc) {                           // chunk order (deterministic); eight loads in flight
      const float4 q = __ldcg(reinterpret_cast<const float4*>(part + (long)cc * part_stride + (long)row * C + col));
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    const float4 r = __ldcg(reinterpret_cast<const float4*>(p2 + (long)row * C + col)), bq = __ldg(reinterpret_cast<const float4*>(b2 + col));
    v[4 * hh] = a.x + bq.x + r.x; v[4 * hh + 1] = a.y + bq.y + r.y; v[4 * hh + 2] = a.z + bq.z + r.z; v[4 * hh + 3] = a.w + bq.w + r.w;
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += v[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.f / C);
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { const float dlt = v[k] - mean; sq = fmaf(dlt, dlt, sq); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.f / C) + LN_EPS);
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int col = hh * 128 + lane * 4;
    const float4 w = __ldg(reinterpret_cast<const float4*>(gw + col)), bq = __ldg(reinterpret_cast<const float4*>(gb + col));
    float4 o;
    o.x = (v[4 * hh] - mean) * rstd * w.x + bq.x; o.y = (v[4 * hh + 1] - mean) * rstd * w.y + bq.y;
    o.z = (v[4 * hh + 2] - mean) * rstd * w.z + bq.z; o.w = (v[4 * hh + 3] - mean) * rstd * w.w + bq.w;
    if (post) {
      const float4 pa = __ldcg(reinterpret_cast<const float4*>(post + (long)row * C + col));
      o.x += pa.x; o.y += pa.y; o.z += pa.z; o.w += pa.w;
    }
    *reinterpret_cast<float4*>(f_out + (long)row * C + col) = o;
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
struct SlotTcWeights {                          // fp16 hi/lo planes [2][Opad][K] per linear layer of one stage
  __half *out_proj, *to_q, *wkT, *wv, *lin1, *lin2, *tw, *cls1, *reg1, *logit;
  __half *tqkv, *tlin1, *tlin2;                 // Video Retriever: q|k|v stacked [2][768][256], FFN [2][TF][256], [2][256][TF]
};
inline void slot_tc_layout(Arena& a, const slotvps_head_desc* d, SlotTcWeights* w) {
  const size_t cc = (size_t)2 * C * C;
  w->out_proj = a.take<__half>(cc); w->to_q = a.take<__half>(cc); w->wkT = a.take<__half>(cc); w->wv = a.take<__half>(cc);
  w->lin1 = a.take<__half>((size_t)2 * d->dim_feedforward * C); w->lin2 = a.take<__half>((size_t)2 * C * d->dim_feedforward);
  w->tw = a.take<__half>(2 * cc); w->cls1 = a.take<__half>(cc); w->reg1 = a.take<__half>(cc);
  w->logit = a.take<__half>((size_t)2 * slot::TILE_N * C);
  w->tqkv = a.take<__half>(3 * cc);
  w->tlin1 = a.take<__half>((size_t)2 * d->temporal_dim_feedforward * C); w->tlin2 = a.take<__half>((size_t)2 * C * d->temporal_dim_feedforward);
}
inline bool slot_tc_supported(const slotvps_head_desc* d) {
  return d->kernel_path == 0 && d->dim_feedforward % 128 == 0 && d->temporal_dim_feedforward % 128 == 0 &&
         d->num_classes <= 32;
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
