// Host-only introspection of the tree compiler (include/dpq.h: dpq_program_*).
#include <cstring>
#include <string>

#include "../../include/dpq.h"
#include "dpq_internal.h"

namespace dpq {
int api_fail(int code, const std::string& msg);
}

struct dpq_program {
    dpq::ScanProgram p;
};

extern "C" {

int dpq_program_compile(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M, int K, int rank,
                        int n_ranks, int chunk_nodes, dpq_program** out) {
    // chunk_nodes < 0 selects the first-generation op program with |chunk_nodes| nodes per chunk
    const int engine = chunk_nodes < 0 ? 1 : 0;
    if (chunk_nodes < 0) chunk_nodes = -chunk_nodes;
    if (!payload || !out) return dpq::api_fail(DPQ_ERR_ARG, "dpq_program_compile: null argument");
    dpq_program* h = new dpq_program();
    std::string err = dpq::compile_program(payload, n_bytes, n_codes, M, K, rank, n_ranks, chunk_nodes, &h->p, engine);
    if (!err.empty()) {
        delete h;
        *out = nullptr;
        return dpq::api_fail(DPQ_ERR_FORMAT, "dpq_program_compile: " + err);
    }
    *out = h;
    return DPQ_OK;
}

int64_t dpq_program_size(dpq_program* h, const char* what) {
    if (!h || !what) return -1;
    std::string w(what);
    const dpq::ScanProgram& p = h->p;
    if (w == "ops") return (int64_t)p.ops.size() * 4;
    if (w == "chunks") return (int64_t)p.chunks.size() * (int64_t)sizeof(dpq::ChunkDesc);
    if (w == "anc") return (int64_t)p.anc.size();
    if (w == "codes") return (int64_t)p.codes.size();
    if (w == "v2") return p.v2 ? 1 : 0;
    if (w == "v2_nf") return p.shape.nf;
    if (w == "v2_lpg") return p.shape.lpg;
    if (w == "cstride") return p.cstride;
    if (w == "n_ops") return (int64_t)p.ops.size();
    if (w == "n_chunks") return (int64_t)p.chunks.size();
    if (w == "n_local") return p.n_local;
    if (w == "base_pos") return p.base_pos;
    if (w == "rb") return p.fmt.rb;
    if (w == "levels") return p.fmt.levels();
    if (w == "n_bytes") return p.local_bytes;
    if (w == "n_diffs") return p.n_diffs;
    return -1;
}

int dpq_program_copy(dpq_program* h, const char* what, void* dst) {
    if (!h || !what || !dst) return dpq::api_fail(DPQ_ERR_ARG, "dpq_program_copy: null argument");
    std::string w(what);
    const dpq::ScanProgram& p = h->p;
    if (w == "ops") memcpy(dst, p.ops.data(), p.ops.size() * 4);
    else if (w == "chunks") memcpy(dst, p.chunks.data(), p.chunks.size() * sizeof(dpq::ChunkDesc));
    else if (w == "anc") memcpy(dst, p.anc.data(), p.anc.size());
    else if (w == "codes") memcpy(dst, p.codes.data(), p.codes.size());
    else return dpq::api_fail(DPQ_ERR_ARG, "dpq_program_copy: unknown array " + w);
    return DPQ_OK;
}

void dpq_program_free(dpq_program* h) { delete h; }

}  // extern "C"
