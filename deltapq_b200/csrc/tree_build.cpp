// DeltaTree build, stages 2-3 (host): edges -> DFS layout -> QNode records -> byte stream.
//
// Reference: create_approx_tree (DCAT.h:970-1065) calls
//   edges_to_tree_index_approx_dfs_layout (DCAT.h:1334-1487; CSR :1067-1104; DFS :1156-1183)
//   qnodes_to_compressed_codes_opt        (DCAT.h:1730-1845)
// with the K x K centroid tables of dmain:101-118 and the table distance of CT.h:827-835.
// This file is the sequential host form (dpq_tree_from_edges: no GPU needed; keeps the
// reference's expression shapes).  dpq_tree_build runs the edge search
// (edges.cu) and the data-parallel form of this stage (layout.cu) on the device; the GPU
// tests compare the two bit for bit.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/dpq.h"
#include "tree_internal.h"

namespace dpq {
int api_fail(int code, const std::string& msg);

// dmain:107-116: float accumulator, each term the double square of the float difference
void centroid_tables(const float* cw, int M, int K, int Ds, std::vector<float>& T) {
    T.assign((size_t)M * K * K, 0.0f);
    for (int m = 0; m < M; ++m)
        for (int a = 0; a < K; ++a) {
            const float* ca = cw + ((size_t)m * K + a) * Ds;
            for (int b = 0; b < K; ++b) {
                const float* cb = cw + ((size_t)m * K + b) * Ds;
                float acc = 0.0f;
                for (int d = 0; d < Ds; ++d) {
                    const float diff = ca[d] - cb[d];
                    acc = (float)((double)acc + (double)diff * (double)diff);
                }
                T[((size_t)m * K + a) * K + b] = acc;
            }
        }
}
}  // namespace dpq

namespace {

struct Layout {
    const uint8_t* codes;
    int M, K;
    int64_t n;
    const std::vector<float>& T;
    // CT.h:827-835: float sum over m ascending
    float pair_dist(uint32_t a, uint32_t b) const {
        const uint8_t* x = codes + (size_t)a * M;
        const uint8_t* y = codes + (size_t)b * M;
        float s = 0.0f;
        for (int m = 0; m < M; ++m) s += T[((size_t)m * K + x[m]) * K + y[m]];
        return s;
    }
};

std::string layout_tree(const uint8_t* codes, int64_t n, int M, int K, const float* cw, int Ds, dpq_tree* t) {
    const int64_t E = n - 1;
    const uint32_t NONE = 0xFFFFFFFFu;
    const std::vector<uint32_t>& edges = t->edges;
    for (int64_t e = 0; e < E; ++e)
        if (edges[2 * e] >= (uint64_t)n || edges[2 * e + 1] >= (uint64_t)n) return "edge endpoint out of range";
    if (t->root >= (uint64_t)n) return "root id out of range";

    // parent of every code id, child lists in emission order (DCAT.h:1067-1104)
    std::vector<uint32_t> parent((size_t)n, NONE), first((size_t)n + 1, 0), kids((size_t)std::max<int64_t>(E, 1));
    for (int64_t e = 0; e < E; ++e) {
        const uint32_t c = edges[2 * e + 1];
        if (c == t->root || parent[c] != NONE) return "edges do not form a tree (a node has two parents)";
        parent[c] = edges[2 * e];
        first[edges[2 * e] + 1]++;
    }
    for (int64_t v = 0; v < n; ++v) first[(size_t)v + 1] += first[(size_t)v];
    {
        std::vector<uint32_t> cur(first.begin(), first.end() - 1);
        for (int64_t e = 0; e < E; ++e) kids[cur[edges[2 * e]]++] = edges[2 * e + 1];
    }

    // farthest descendant within 16 levels, per node and per (node, child branch) (DCAT.h:1396-1417)
    std::vector<float> T;
    dpq::centroid_tables(cw, M, K, Ds, T);
    Layout L{codes, M, K, n, T};
    std::vector<float> far((size_t)n, 0.0f), far_via((size_t)n, 0.0f);
    for (int64_t v = 0; v < n; ++v) {
        uint32_t below = (uint32_t)v;
        uint32_t anc = parent[(size_t)v];
        for (int hop = 0; anc != NONE && hop < 16; ++hop) {
            const float d = L.pair_dist((uint32_t)v, anc);
            if (d > far[anc]) far[anc] = d;
            if (d > far_via[below]) far_via[below] = d;
            below = anc;
            anc = parent[anc];
        }
    }
    // children by far_via descending, stable (DCAT.h:1421-1426)
    for (int64_t v = 0; v < n; ++v) {
        uint32_t* b = kids.data() + first[(size_t)v];
        uint32_t* e = kids.data() + first[(size_t)v + 1];
        if (e - b > 1) std::stable_sort(b, e, [&](uint32_t x, uint32_t y) { return far_via[x] > far_via[y]; });
    }

    // pre-order numbering from the root (DCAT.h:1156-1183)
    t->vec_id.assign((size_t)n, 0);
    t->parent_pos.assign((size_t)n, NONE);
    t->child_num.assign((size_t)n, 0);
    t->depth.assign((size_t)n, 0);
    struct Frame {
        uint32_t vid, next, pos;
    };
    std::vector<Frame> path;
    path.push_back({t->root, first[t->root], 0});
    t->vec_id[0] = t->root;
    int64_t placed = 1;
    while (!path.empty()) {
        Frame& f = path.back();
        if (f.next == first[(size_t)f.vid + 1]) {  // subtree complete: descendants = placed - pos - 1
            t->child_num[f.pos] = (uint32_t)(placed - 1 - f.pos);
            path.pop_back();
            continue;
        }
        const uint32_t child = kids[f.next++];
        const uint32_t pos = (uint32_t)placed++;
        if (placed > n) return "edges do not form a tree (cycle or duplicate child)";
        t->vec_id[pos] = child;
        t->parent_pos[pos] = f.pos;
        if (path.size() > 255) return "tree deeper than 255 levels";
        t->depth[pos] = (uint8_t)path.size();
        path.push_back({child, first[child], pos});
    }
    if (placed != n) return "edges do not span all codes from the root";
    t->max_dist.resize((size_t)n);
    t->max_dist2p.resize((size_t)n);
    for (int64_t p = 0; p < n; ++p) {
        t->max_dist[(size_t)p] = std::sqrt(far[t->vec_id[(size_t)p]]);
        t->max_dist2p[(size_t)p] = std::sqrt(far_via[t->vec_id[(size_t)p]]);
    }
    t->codes_by_pos.resize((size_t)n * M);
    for (int64_t p = 0; p < n; ++p)
        memcpy(&t->codes_by_pos[(size_t)p * M], codes + (size_t)t->vec_id[(size_t)p] * M, (size_t)M);
    return std::string();
}

// DCAT.h:1765-1842 (M == 8) / the extension format for M > 8: ceil(M/8) bitmap
// bytes little-endian.  Nodes in position order, two depth nibbles per pair.
std::string write_stream(dpq_tree* t) {
    const int M = t->M;
    const int64_t n = t->n;
    const int bmb = (M + 7) / 8;
    const uint8_t* cp = t->codes_by_pos.data();
    int64_t nd = 0;
    for (int64_t p = 1; p < n; ++p) {
        const uint8_t* c = cp + (size_t)p * M;
        const uint8_t* q = cp + (size_t)t->parent_pos[(size_t)p] * M;
        for (int m = 0; m < M; ++m) nd += c[m] != q[m];
        // M <= 8: the reference reader masks every depth nibble with &7 (DCAT.h:3794), so a deeper
        // tree (max_height_folds >= 2) cannot be represented; the M > 8 extension keeps 4 bits
        if (t->depth[(size_t)p] > (M > 8 ? 15 : 7))
            return M > 8 ? "depth > 15 does not fit the stream's depth nibble"
                         : "depth > 7 is not representable when M <= 8 (the reader masks depth nibbles with &7, DCAT.h:3794)";
    }
    t->n_diffs = nd;
    const int64_t total = (int64_t)M + nd + (int64_t)bmb * (n - 1) + n / 2;  // M = 8: 8 + n_diffs + (3(n-1)+1)/2
    t->payload.assign((size_t)total, 0);
    uint8_t* out = t->payload.data();
    int64_t o = 0;
    memcpy(out, cp, (size_t)M);
    o = M;
    for (int64_t p = 1; p < n; ++p) {
        if (p & 1) out[o++] = (uint8_t)(t->depth[(size_t)p] | (p + 1 < n ? t->depth[(size_t)p + 1] << 4 : 0));
        const uint8_t* c = cp + (size_t)p * M;
        const uint8_t* q = cp + (size_t)t->parent_pos[(size_t)p] * M;
        uint32_t bm = 0;
        for (int m = 0; m < M; ++m) bm |= (uint32_t)(c[m] != q[m]) << m;
        for (int b = 0; b < bmb; ++b) out[o++] = (uint8_t)(bm >> (8 * b));
        for (int m = 0; m < M; ++m)
            if ((bm >> m) & 1u) out[o++] = c[m];
    }
    if (o != total) return "stream size mismatch";
    return std::string();
}

// QNode records as the reference writes them (DCAT.h:79-101, 1437-1445, 1484): 60 bytes each,
// N + 1 entries, M == 8 only (array<Diff, 8>).
void write_qnodes8(const dpq_tree* t, uint8_t* dst) {
    const int64_t n = t->n;
    memset(dst, 0, (size_t)(n + 1) * 60);
    const uint8_t* cp = t->codes_by_pos.data();
    for (int64_t p = 0; p <= n; ++p) {
        uint8_t* r = dst + (size_t)p * 60;
        const uint32_t one = 1;
        memcpy(r + 16, &one, 4);  // sub_tree_size
        if (p == n) break;
        const uint32_t cps = (uint32_t)p + 1;
        memcpy(r + 0, &t->vec_id[(size_t)p], 4);
        memcpy(r + 4, &t->parent_pos[(size_t)p], 4);
        memcpy(r + 8, &cps, 4);
        memcpy(r + 12, &t->child_num[(size_t)p], 4);
        memcpy(r + 24, &t->max_dist[(size_t)p], 4);
        memcpy(r + 28, &t->max_dist2p[(size_t)p], 4);
        r[33] = t->depth[(size_t)p];
        const uint8_t* c = cp + (size_t)p * 8;
        int nd = 0;
        if (p == 0) {
            for (int m = 0; m < 8; ++m) {
                r[34 + 3 * m] = (uint8_t)m;
                r[35 + 3 * m] = 0xFF;
                r[36 + 3 * m] = c[m];
            }
            nd = 8;
        } else {
            const uint8_t* q = cp + (size_t)t->parent_pos[(size_t)p] * 8;
            for (int m = 0; m < 8; ++m)
                if (c[m] != q[m]) {
                    r[34 + 3 * nd] = (uint8_t)m;
                    r[35 + 3 * nd] = q[m];
                    r[36 + 3 * nd] = c[m];
                    ++nd;
                }
        }
        r[32] = (uint8_t)nd;
    }
}

int finish_tree(const uint8_t* codes, int64_t n, int M, int K, const float* cw, int Ds, dpq_tree* t, dpq_tree** out) {
    std::string err = layout_tree(codes, n, M, K, cw, Ds, t);
    if (err.empty()) err = write_stream(t);
    if (!err.empty()) {
        delete t;
        return dpq::api_fail(DPQ_ERR_FORMAT, "dpq_tree: " + err);
    }
    *out = t;
    return DPQ_OK;
}

bool bad_args(const uint8_t* codes, int64_t n, int M, int K, const float* cw, int Ds, dpq_tree** out) {
    return !codes || !cw || !out || n < 1 || n >= 0x7FFFFFFFLL || M < 1 || M > 16 || K < 1 || K > 256 || Ds < 1;
}

}  // namespace

extern "C" {

int dpq_tree_from_edges(const uint8_t* codes, int64_t n_codes, int M, int K, const float* codewords, int Ds,
                        const uint32_t* edges, uint32_t root_id, dpq_tree** out) {
    if (bad_args(codes, n_codes, M, K, codewords, Ds, out) || (n_codes > 1 && !edges))
        return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_from_edges: bad argument");
    *out = nullptr;
    if (K < 256)
        for (int64_t i = 0; i < n_codes * M; ++i)
            if (codes[i] >= K) return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_from_edges: codes hold a centroid id >= K");
    dpq_tree* t = new dpq_tree();
    t->M = M;
    t->K = K;
    t->n = n_codes;
    t->root = root_id;
    if (n_codes > 1) t->edges.assign(edges, edges + 2 * (n_codes - 1));
    return finish_tree(codes, n_codes, M, K, codewords, Ds, t, out);
}

int dpq_tree_build(const uint8_t* codes, int64_t n_codes, int M, int K, const float* codewords, int Ds,
                   int max_height_folds, int method, dpq_tree** out) {
    if (bad_args(codes, n_codes, M, K, codewords, Ds, out))
        return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_build: bad argument");
    *out = nullptr;
    if (int rc0 = dpq::check_code_range(codes, n_codes, M, K)) return rc0;
    dpq_tree* t = new dpq_tree();
    t->M = M;
    t->K = K;
    t->n = n_codes;
    t->edges.assign((size_t)std::max<int64_t>(2 * (n_codes - 1), 2), 0);
    const auto t0 = std::chrono::steady_clock::now();
    int rc = dpq_find_edges(codes, n_codes, M, K, max_height_folds, method, t->edges.data(), &t->root);
    const auto t1 = std::chrono::steady_clock::now();
    t->edge_us = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
    if (rc) {
        delete t;
        return rc;
    }
    t->edges.resize((size_t)(2 * (n_codes - 1)));
    // layout + stream on the device (layout.cu); DPQ_HOST_LAYOUT=1 keeps the sequential host walk
    const char* host_layout = getenv("DPQ_HOST_LAYOUT");
    if (host_layout && host_layout[0] == '1') return finish_tree(codes, n_codes, M, K, codewords, Ds, t, out);
    rc = dpq::layout_tree_device(codes, n_codes, M, K, codewords, Ds, t);
    t->layout_us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t1).count();
    if (rc) {
        delete t;
        return rc;
    }
    *out = t;
    return DPQ_OK;
}

int dpq_tree_build_device(const uint8_t* codes, int64_t n_codes, int M, int K, const float* codewords, int Ds,
                          int max_height_folds, int method, dpq_tree** out) {
    if (bad_args(codes, n_codes, M, K, codewords, Ds, out))
        return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_build_device: bad argument");
    *out = nullptr;
    if (int rc0 = dpq::check_code_range(codes, n_codes, M, K)) return rc0;
    dpq_tree* t = new dpq_tree();
    t->M = M;
    t->K = K;
    t->n = n_codes;
    uint32_t* d_edges = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = dpq::find_edges_device(codes, n_codes, M, K, max_height_folds, method, &d_edges, &t->root);
    const auto t1 = std::chrono::steady_clock::now();
    t->edge_us = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
    if (!rc) rc = dpq::layout_tree_device(codes, n_codes, M, K, codewords, Ds, t, d_edges, true);
    if (d_edges) dpq_free(d_edges);
    t->layout_us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t1).count();
    if (rc) {
        delete t;
        return rc;
    }
    *out = t;
    return DPQ_OK;
}

int64_t dpq_tree_size(dpq_tree* t, const char* what) {
    if (!t || !what) return -1;
    const std::string w(what);
    if (t->on_device) {  // what a device-resident tree can still hand out (copied D2H on request)
        if (w == "vec_id") return t->n * 4;
        if (w == "depth") return t->n;
        if (w == "codes_by_pos") return t->n * t->M;
        if (w == "payload") return t->payload_bytes;
        if (w == "on_device") return 1;
        if (w.rfind("depth_hist_", 0) == 0) {
            const size_t d = (size_t)atoi(w.c_str() + 11);
            return d < t->depth_hist.size() ? t->depth_hist[d] : 0;
        }
        if (w == "edges" || w == "parent_pos" || w == "child_num" || w == "max_dist" || w == "max_dist2p" || w == "qnodes")
            return -1;
    }
    if (w == "edges") return (int64_t)t->edges.size() * 4;
    if (w == "vec_id" || w == "parent_pos" || w == "child_num" || w == "max_dist" || w == "max_dist2p")
        return t->n * 4;
    if (w == "depth") return t->n;
    if (w == "codes_by_pos") return (int64_t)t->codes_by_pos.size();
    if (w == "payload") return (int64_t)t->payload.size();
    if (w == "qnodes") return t->M == 8 ? (t->n + 1) * 60 : -1;
    if (w == "root_id") return t->root;
    if (w == "n_diffs") return t->n_diffs;
    if (w == "n_codes") return t->n;
    if (w == "edge_us") return t->edge_us;
    if (w == "layout_us") return t->layout_us;
    return -1;
}

int dpq_tree_copy(dpq_tree* t, const char* what, void* dst) {
    if (!t || !what || !dst) return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_copy: null argument");
    const std::string w(what);
    if (t->on_device) {
        if (w == "vec_id") return dpq::tree_copy_device(t, t->d_vec_id, dst, (size_t)t->n * 4);
        if (w == "depth") return dpq::tree_copy_device(t, t->d_depth, dst, (size_t)t->n);
        if (w == "codes_by_pos") return dpq::tree_copy_device(t, t->d_codes_by_pos, dst, (size_t)t->n * t->M);
        if (w == "payload") return dpq::tree_copy_device(t, t->d_payload, dst, (size_t)t->payload_bytes);
        return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_copy: a device-resident tree keeps vec_id, depth, codes_by_pos and payload only");
    }
    auto put = [&](const void* src, size_t bytes) {
        if (bytes) memcpy(dst, src, bytes);
        return DPQ_OK;
    };
    if (w == "edges") return put(t->edges.data(), t->edges.size() * 4);
    if (w == "vec_id") return put(t->vec_id.data(), t->vec_id.size() * 4);
    if (w == "parent_pos") return put(t->parent_pos.data(), t->parent_pos.size() * 4);
    if (w == "child_num") return put(t->child_num.data(), t->child_num.size() * 4);
    if (w == "max_dist") return put(t->max_dist.data(), t->max_dist.size() * 4);
    if (w == "max_dist2p") return put(t->max_dist2p.data(), t->max_dist2p.size() * 4);
    if (w == "depth") return put(t->depth.data(), t->depth.size());
    if (w == "codes_by_pos") return put(t->codes_by_pos.data(), t->codes_by_pos.size());
    if (w == "payload") return put(t->payload.data(), t->payload.size());
    if (w == "qnodes") {
        if (t->M != 8) return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_copy: QNode records exist for M == 8 only");
        write_qnodes8(t, reinterpret_cast<uint8_t*>(dst));
        return DPQ_OK;
    }
    return dpq::api_fail(DPQ_ERR_ARG, "dpq_tree_copy: unknown array " + w);
}

void dpq_tree_free(dpq_tree* t) { delete t; }

}  // extern "C"
