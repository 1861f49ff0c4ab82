// placeholder until the tcgen05 kernel lands (same translation unit name)
#include "tq_common.cuh"
extern "C" {
size_t tq_linear_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }
int tq_linear_qdq_bf16(const void*, const void*, const float*, float*, void*, int64_t, int64_t, int64_t, int32_t,
                       const float*, const float*, int32_t, int32_t, tq_qspec, int64_t, float*, void*, size_t,
                       void*) { return TQ_EUNSUPPORTED; }
int tq_split3_bf16(const float*, void*, int64_t, int64_t, void*) { return TQ_EUNSUPPORTED; }
}
