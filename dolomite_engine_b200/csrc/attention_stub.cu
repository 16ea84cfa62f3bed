// placeholder until attention kernels land (same translation unit name will be replaced)
#include "common.cuh"
#include "../../include/dolomite_b200.h"
extern "C" int dolomite_b200_attn_varlen_fwd(const void*, int64_t, void*, float*, const int32_t*, int, int64_t, int, int,
                                             int, int, float, void*) {
    return dolo_set_error("attn_varlen_fwd: not built");
}
extern "C" int64_t dolomite_b200_attn_varlen_bwd_workspace_bytes(int64_t, int, int, int) { return 0; }
extern "C" int dolomite_b200_attn_varlen_bwd(const void*, const void*, int64_t, const void*, const float*, void*,
                                             const int32_t*, int, int64_t, int, int, int, int, float, void*, void*) {
    return dolo_set_error("attn_varlen_bwd: not built");
}
