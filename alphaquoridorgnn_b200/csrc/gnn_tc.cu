// bf16 tcgen05 tensor-core path of the GCN forward (placeholder until the UMMA kernel lands).
#include "aq_common.cuh"
int aq_gcn_forward_tc(const float *, const AqState *, int64_t, float *, cudaStream_t) {
    return aq_set_error(AQ_ERR_UNSUPPORTED, "bf16 tcgen05 path not built");
}
