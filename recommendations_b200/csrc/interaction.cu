// placeholder: replaced by the tcgen05 kernel
#include "common.cuh"
using namespace recemb;
extern "C" int recemb_dot_interaction_fwd(const void*, int64_t, int32_t, int32_t, void*, int, recemb_stream_t) {
  set_error("dot_interaction_fwd: not built yet");
  return RECEMB_ERR_UNSUPPORTED;
}
extern "C" int recemb_dot_interaction_bwd(const void*, const void*, int64_t, int32_t, int32_t, void*, int, recemb_stream_t) {
  set_error("dot_interaction_bwd: not built yet");
  return RECEMB_ERR_UNSUPPORTED;
}
