// stubs.cu -- entry points declared in include/monovo_b200.h whose kernels are not built yet.
// They fail loudly with MVO_ERR_UNSUPPORTED (never a CPU fallback).
#include "context.cuh"
#define MVO_STUB(c, name)                                  \
  do {                                                     \
    if (c) (c)->set_error(name ": not implemented yet");   \
    return MVO_ERR_UNSUPPORTED;                            \
  } while (0)
extern "C" {
int mvo_find_homography(mvo_ctx* c, const float*, const float*, int, double, double*, uint8_t*, int*) { MVO_STUB(c, "mvo_find_homography"); }
int mvo_find_fundamental(mvo_ctx* c, const float*, const float*, int, double, double, double*, uint8_t*, int*) { MVO_STUB(c, "mvo_find_fundamental"); }
int mvo_find_essential(mvo_ctx* c, const float*, const float*, int, const double*, double, double, double*, uint8_t*, int*) { MVO_STUB(c, "mvo_find_essential"); }
int mvo_recover_pose(mvo_ctx* c, const double*, const float*, const float*, int, const double*, double*, double*, uint8_t*, int*) { MVO_STUB(c, "mvo_recover_pose"); }
int mvo_triangulate(mvo_ctx* c, const double*, const double*, const float*, const float*, int, float*) { MVO_STUB(c, "mvo_triangulate"); }
int mvo_score_hypotheses(mvo_ctx* c, int, const float*, const float*, int, const double*, double, int, int32_t*, int32_t*, double*) { MVO_STUB(c, "mvo_score_hypotheses"); }
int mvo_group_step(mvo_ctx* c, const uint8_t*, int, int, int, int, const double*, mvo_frame_result*) { MVO_STUB(c, "mvo_group_step"); }
int mvo_stage_ms(mvo_ctx* c, const char*, float*) { MVO_STUB(c, "mvo_stage_ms"); }
}
