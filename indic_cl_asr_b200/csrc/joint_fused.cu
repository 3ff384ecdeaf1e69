// joint_fused.cu — placeholder until the tcgen05 kernel lands (entry points exist so the ABI is complete).
#include "common.cuh"
using namespace clasr;
extern "C" size_t clasr_joint_workspace_bytes(int B, int T, int U1, int H, int Vp, int precision) { return 0; }
extern "C" int clasr_joint_rnnt_fwd(const float*, const float*, const float*, const float*, const int64_t*,
                                    const int64_t*, const int64_t*, int, int, int, int, int, int, int, int, float,
                                    float*, float*, void*, size_t, void*) {
  set_error("joint_rnnt_fwd: not implemented");
  return CLASR_STATUS_INVALID_VALUE;
}
extern "C" int clasr_joint_rnnt_bwd(const float*, const float*, const float*, const float*, const int64_t*,
                                    const int64_t*, const int64_t*, int, int, int, int, int, int, int, int, float,
                                    float, const float*, float*, float*, float*, float*, void*, size_t, void*) {
  set_error("joint_rnnt_bwd: not implemented");
  return CLASR_STATUS_INVALID_VALUE;
}
