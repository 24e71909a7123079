// stubs.cu -- entry points declared in include/monovo_b200.h whose kernels are not built yet.
// They fail loudly with MVO_ERR_UNSUPPORTED (never a CPU fallback).
#include "context.cuh"
#define MVO_STUB(c, name)                                  \
  do {                                                     \
    if (c) (c)->set_error(name ": not implemented yet");   \
    return MVO_ERR_UNSUPPORTED;                            \
  } while (0)
extern "C" {
int mvo_group_step(mvo_ctx* c, const uint8_t*, int, int, int, int, const double*, mvo_frame_result*) { MVO_STUB(c, "mvo_group_step"); }
int mvo_stage_ms(mvo_ctx* c, const char*, float*) { MVO_STUB(c, "mvo_stage_ms"); }
}
