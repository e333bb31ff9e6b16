// tcgen05 / TMEM / TMA implementation of the pixel-side retriever kernels (kernel_path = 0).
// (placeholder while the fp32 path is brought up: tc_supported() == false routes everything
//  to pixel_fp32.cuh)
#pragma once
#include "common.cuh"

namespace slotvps {

struct TcStageOperands { void* base = nullptr; };
struct TcWorkspace { void* base = nullptr; };

inline void tc_stage_layout(Arena& a, TcStageOperands* o) { (void)a; (void)o; }
inline void tc_workspace_layout(Arena& a, const slotvps_head_desc* d, TcWorkspace* w) { (void)a; (void)d; (void)w; }
inline int tc_prepare_stage(const slotvps_stage_params& sp, const float* Wk_c, const float* bk_c, const float* Wv_c,
                            const float* bv_c, TcStageOperands& o, cudaStream_t s) {
  (void)sp; (void)Wk_c; (void)bk_c; (void)Wv_c; (void)bv_c; (void)o; (void)s;
  return SLOTVPS_OK;
}
inline bool tc_supported(const slotvps_head_desc* d, int level) { (void)d; (void)level; return false; }
inline int pixel_attention_tc(const float* x, long x_bs, const float* pos, long pos_bs, const TcStageOperands& ops,
                              const TcWorkspace& ws, const float* G, const float* g0, const float* g1, float* Z, float* a0,
                              float* a1, int T, int N, int h, int w, cudaStream_t s) {
  (void)x; (void)x_bs; (void)pos; (void)pos_bs; (void)ops; (void)ws; (void)G; (void)g0; (void)g1; (void)Z; (void)a0; (void)a1;
  (void)T; (void)N; (void)h; (void)w; (void)s;
  return fail(SLOTVPS_EUNSUPPORTED, "tensor-core path not built%s%s");
}

}  // namespace slotvps
