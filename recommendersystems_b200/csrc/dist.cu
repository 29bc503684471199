// dist.cu -- row-partitioned mode (K10): placeholder entry points until the NCCL path lands.
#include "graph.h"

struct rwr_comm {
    int rank = 0, n_ranks = 1;
};

extern "C" {

int rwr_comm_unique_id(void* id128) {
    (void)id128;
    rwr_set_error("row-partitioned mode is not built yet");
    return RWR_E_UNSUPPORTED;
}
int rwr_comm_create(int32_t, int32_t, const void*, const rwr_opts*, rwr_comm** out) {
    if (out) *out = nullptr;
    rwr_set_error("row-partitioned mode is not built yet");
    return RWR_E_UNSUPPORTED;
}
void rwr_comm_destroy(rwr_comm* c) { delete c; }
int rwr_synth_create_partitioned(const rwr_synth_spec*, const rwr_opts*, rwr_comm*, rwr_graph** out) {
    if (out) *out = nullptr;
    rwr_set_error("row-partitioned mode is not built yet");
    return RWR_E_UNSUPPORTED;
}
int rwr_graph_create_partitioned(int32_t, const int64_t*, const int32_t*, int64_t, const int32_t*, const int32_t*,
                                 const int32_t*, const double*, const rwr_opts*, rwr_comm*, rwr_graph** out) {
    if (out) *out = nullptr;
    rwr_set_error("row-partitioned mode is not built yet");
    return RWR_E_UNSUPPORTED;
}

}  // extern "C"
