// placeholder until the prefix-LM attention kernels land (replaced in the next commit)
#include "host_utils.h"
#include "llamax_b200.h"
extern "C" {
int llamax_attn_fwd(const void*, int64_t, const void*, int64_t, const void*, int64_t, void*, int64_t, void*, int64_t,
                    int64_t, int32_t, int32_t, int32_t, int64_t, float, void*) {
  return lx::set_error(LLAMAX_ERR_ARG, "attn_fwd: not built yet");
}
int llamax_attn_bwd(const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*, int64_t, const void*,
                    const void*, int64_t, void*, int64_t, void*, int64_t, void*, int64_t, void*, void*, int64_t, int64_t,
                    int32_t, int32_t, int32_t, int64_t, float, void*) {
  return lx::set_error(LLAMAX_ERR_ARG, "attn_bwd: not built yet");
}
}
