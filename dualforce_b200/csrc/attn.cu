// placeholder until the tcgen05 attention kernel lands
#include "host_utils.h"
#include "../../include/mova_b200.h"
extern "C" int mova_b200_attn_fwd(const void*, int64_t, int64_t, const void*, int64_t, int64_t, const void*, int64_t,
                                  int64_t, void*, int64_t, int64_t, float*, int, int, int, int, int, float, void*) {
  mv::set_error("mova_b200_attn_fwd: not built yet");
  return -1;
}
