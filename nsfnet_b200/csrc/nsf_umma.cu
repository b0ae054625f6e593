// tcgen05 (UMMA) path -- placeholder until the 3xTF32 kernels land.
#include "nsf_internal.h"

extern "C" int nsf_selftest_umma(int, int32_t, const float*, const float*, float*, int32_t, int32_t, void*) {
  nsf_set_error("nsf_selftest_umma: not built yet");
  return NSF_E_SHAPE;
}
