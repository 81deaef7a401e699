// placeholder until the tcgen05 kernel lands
#include "common.cuh"
#include "resattn.h"
bool resattn_tc_supported(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t) { return false; }
int resattn_fwd_tc(const void*, int64_t, const void*, int64_t, const void*, int64_t, const float*,
                   int64_t, const void*, const float*, void*, void*, int64_t, float*, int64_t,
                   int64_t, int64_t, int64_t, int64_t, cudaStream_t) { return MMEMO_ERR_SHAPE; }
