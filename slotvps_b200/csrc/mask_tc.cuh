// Mask-logit projection (generate_final_outputs, vps_temporal_slots.py:145-154) on the tensor pipe, reading
// the finest level's fp16 operand planes that the head left in its workspace (no second pass over the
// fp32 feature for the contraction):
//   M[n,p] = fg_scale * ((sum_c e'[n,c] x[c,p] + d[n]) * rn[p]) + fg_shift
// e' = emb * feat_bn scale, d = emb . feat_bn shift, rn = 1 / max(||feat_bn(x_p)||, 1e-12) (feat_rnorm_kernel).
// GEMM: M = 128 pixels (TMEM lanes), N = 112 (slots, <= 104 held), K = 256, fp16 hi/lo x3; HBM-bound:
// reads the x planes (1 KB/pixel) and writes N fp32 per pixel.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "pixel_tc.cuh"

namespace slotvps {
namespace mask {
constexpr int TILE_M = 128;
constexpr int NROW = attn::NROW, NPAD = attn::NPAD;
constexpr int SLOT_BYTES = 16384, NSLOT = 7;
constexpr int E_SUB = NROW * 128;
constexpr int E_BYTES = 2 * 4 * E_SUB;
constexpr int OFF_E = NSLOT * SLOT_BYTES;
constexpr int OFF_MISC = OFF_E + E_BYTES;
constexpr int SMEM_BYTES = OFF_MISC + 2048 + 1024;
constexpr int THREADS = 320;                   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant: 64 / 48 slot columns each)
constexpr uint32_t IDESC = tc::make_idesc_f16(128, NPAD, 0, 0);
}  // namespace mask

__global__ void __launch_bounds__(mask::THREADS, 1)
mask_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_e, const float* __restrict__ dn,
               const float* __restrict__ rn, const float* __restrict__ aff, float* __restrict__ out, int N, int P, int row0, int lo_row,
               int rn_is_ss) {
  using namespace mask;
  extern __shared__ uint8_t raw_smem[];
  const uint32_t raw = tc::smem_u32(raw_smem);
  uint8_t* smem = raw_smem + ((1024 - (raw & 1023)) & 1023);
  uint8_t* misc = smem + OFF_MISC;
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);      // [NSLOT]
  uint64_t* empty = full + NSLOT;                          // [NSLOT]
  uint64_t* tfull = empty + NSLOT;                         // [2]
  uint64_t* tempty = tfull + 2;                            // [2]
  uint64_t* efull = tempty + 2;                            // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(efull + 1);
  float* dsm = reinterpret_cast<float*>(misc + 256);       // [NPAD]
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // provably warp-uniform
  const int n_tiles = (P + TILE_M - 1) / TILE_M;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmap_x);
    tc::tma_prefetch_desc(&tmap_e);
    for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 256); }
    tc::mbar_init(efull, 1);
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < NPAD; i += THREADS) dsm[i] = i < N ? dn[i] : 0.f;
  if (warp == 1) { tc::tmem_alloc(tmem_ptr, 256); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_expect_tx(efull, E_BYTES);
      for (int pl = 0; pl < 2; ++pl)
        for (int ks = 0; ks < 4; ++ks) tc::tma_load_2d(smem + OFF_E + (pl * 4 + ks) * E_SUB, &tmap_e, ks * 64, pl * NROW, efull);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row = row0 + tile * TILE_M;
        for (int ks = 0; ks < 4; ++ks)
          for (int pl = 0; pl < 2; ++pl, ++it) {
            const int s = it % NSLOT;
            tc::mbar_wait(&empty[s], ((it / NSLOT) & 1) ^ 1);
            tc::mbar_expect_tx(&full[s], SLOT_BYTES);
            tc::tma_load_2d(smem + s * SLOT_BYTES, &tmap_x, 0, (pl * 4 + ks) * lo_row + row, &full[s]);
          }
      }
    }
  } else if (warp == 1) {
    {                                                         // all lanes: warp-uniform issue loop, one elected lane issues (tc::elect_one)
      const bool el = tc::elect_one();
      tc::mbar_wait(efull, 0);
      tc::tc_fence_after();
      const uint32_t e_base = tc::smem_u32(smem + OFF_E);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const int g = ti & 1, u = ti >> 1;
        tc::mbar_wait(&tempty[g], (u & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d = tmem_base + g * 128;
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t deh = tc::make_smem_desc_sw128(e_base + ks * E_SUB, 16, 1024);
          const uint64_t del = tc::make_smem_desc_sw128(e_base + (4 + ks) * E_SUB, 16, 1024);
          for (int pl = 0; pl < 2; ++pl, ++it) {
            const int s = it % NSLOT;
            tc::mbar_wait(&full[s], (it / NSLOT) & 1);
            tc::tc_fence_after();
            const uint64_t da = tc::make_smem_desc_sw128(tc::smem_u32(smem + s * SLOT_BYTES), 16, 1024);
            if (el) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                tc::umma_bf16(d, da + 2 * k, deh + 2 * k, IDESC, (ks | pl | k) != 0);
                if (pl == 0) tc::umma_bf16(d, da + 2 * k, del + 2 * k, IDESC, 1);
              }
              tc::umma_commit(&empty[s]);
            }
          }
        }
        if (el) tc::umma_commit(&tfull[g]);
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane, jh = (warp - 2) >> 2;    // jh: which half of the slot columns this warp stores
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float sg = aff[0], tg = aff[1];
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int g = ti & 1, u = ti >> 1;
      const int p = tile * TILE_M + r;
      const bool pv = p < P;
      // rn: 1 / max(||feat_bn(x_p)||, 1e-12), or (rn_is_ss = stride between them) the four partial squared norms left by
      // the level-fusion epilogue, added in a fixed order
      float rnp = pv ? __ldg(rn + p) : 1.f;
      if (rn_is_ss && pv) rnp = (rnp + __ldg(rn + rn_is_ss + p)) + (__ldg(rn + 2L * rn_is_ss + p) + __ldg(rn + 3L * rn_is_ss + p));
      const float scale = pv ? (rn_is_ss ? 1.f / fmaxf(sqrtf(rnp), 1e-12f) : rnp) * sg : 0.f;
      tc::mbar_wait(&tfull[g], u & 1);
      tc::tc_fence_after();
#pragma unroll 1
      for (int j = 2 * jh; j < 2 * jh + 2; ++j) {
        float v[32];
        tc::tmem_ld32(tmem_base + lane_addr + g * 128 + j * 32, v);
        tc::tmem_ld_wait();
        if (j == 2 * jh + 1) { tc::tc_fence_before(); tc::mbar_arrive(&tempty[g]); }
        if (!pv) continue;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int n = j * 32 + c;
          if (n < N) out[(long)n * P + p] = fmaf(v[c] + dsm[n], scale, tg);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 256); }
}

// planes: the level's x planes (hi at row 0.., lo at lo_row..), frame rows start at row0; eplanes [2][104][256]
inline int mask_tc_launch(const __half* planes, long plane_rows_total, long lo_row, long row0, const __half* eplanes, const float* dn,
                          const float* rn, const float* aff, float* out, int N, int P, cudaStream_t s, int rn_is_ss = 0) {
  CUtensorMap mx, me;
  SV_TRY(tc::make_tmap_h16_sw128(&mx, planes, (uint64_t)4 * plane_rows_total, 64, mask::TILE_M));      // ks-major sub-planes [2][4][rows][64]
  SV_TRY(tc::make_tmap_h16_sw128(&me, eplanes, (uint64_t)2 * mask::NROW, C, mask::NROW));
  SV_TRY(ensure_dyn_smem((const void*)mask_tc_kernel, mask::SMEM_BYTES));
  const int n_tiles = ceil_div(P, mask::TILE_M);
  const int grid = n_tiles < main_ctas() ? n_tiles : main_ctas();
  g_prof_grid = grid;
  mask_tc_kernel<<<grid, mask::THREADS, mask::SMEM_BYTES, s>>>(mx, me, dn, rn, aff, out, N, P, (int)row0, (int)lo_row, rn_is_ss);
  SV_CHECK_LAUNCH("mask_tc");
  return SLOTVPS_OK;
}

}  // namespace slotvps
