// DeltaTree build: edge search ("find_edge", DCAT.h:1207-1332 driver, :445-627 round).
#include <cuda_runtime.h>

#include <string>

#include "../../include/dpq.h"

namespace dpq {
int api_fail(int code, const std::string& msg);
int api_check_device();
int api_device();
}  // namespace dpq

extern "C" int dpq_find_edges(const uint8_t* codes, int64_t n_codes, int M, int K, int max_height_folds,
                              int method, uint32_t* edges, uint32_t* root_id) {
    (void)codes; (void)n_codes; (void)M; (void)K; (void)max_height_folds; (void)method; (void)edges; (void)root_id;
    return dpq::api_fail(DPQ_ERR_ARG, "dpq_find_edges: not built yet");
}
