// dpq_tree: the result object of dpq_tree_build / dpq_tree_from_edges (include/dpq.h), shared by
// the host layout (tree_build.cpp) and the device layout (layout.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct dpq_tree {
    int M = 0, K = 0;
    int64_t n = 0;
    uint32_t root = 0;
    int64_t n_diffs = 0;
    int64_t edge_us = 0, layout_us = 0;  // wall time of the two build stages (dpq_tree_build)
    std::vector<uint32_t> edges;  // [n-1][2]
    std::vector<uint32_t> vec_id, parent_pos, child_num;
    std::vector<uint8_t> depth;
    std::vector<float> max_dist, max_dist2p;
    std::vector<uint8_t> payload;
    std::vector<uint8_t> codes_by_pos;  // [n][M]
};

namespace dpq {
// dmain:107-116: the K x K centroid distance tables that order the children, [M][K][K]
void centroid_tables(const float* cw, int M, int K, int Ds, std::vector<float>& T);
// DFS layout + stream on the GPU (layout.cu): fills every array of *t from t->edges / t->root.
// codes: [n][M] host pointer.  Returns 0 or a DPQ_ERR_* code (text via dpq_last_error).
int layout_tree_device(const uint8_t* codes, int64_t n, int M, int K, const float* cw, int Ds, dpq_tree* t);
}  // namespace dpq
