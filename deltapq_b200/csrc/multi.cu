// Multi-GPU search behind the C ABI (include/dpq.h: dpq_multi_*): one process, one host thread,
// one DeltaTree shard per GPU (whole depth-1 subtrees), local top-k on every GPU,
// ONE collective -- an NCCL all-gather of the Q x k result keys over NVLink -- and a k-way merge
// kernel.  This is what the drop-in `deltapq -task query -gpus N` uses; bench.py does the same
// with torch.distributed as the NCCL plumbing.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"), so libdpq.so has no link-time dependency on
// it and single-GPU users never load it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dpq.h"

namespace dpq {
int api_fail(int code, const std::string& msg);
}

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;  // 0 = ncclSuccess
constexpr int kNcclUint64 = 5;  // ncclUint64 in ncclDataType_t
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string load() {
        if (lib) return std::string();
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return std::string("cannot load libnccl.so.2: ") + dlerror();
#define DPQ_SYM(field, sym)                                                  \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, sym));              \
    if (!field) return std::string("libnccl lacks ") + sym;
        DPQ_SYM(CommInitAll, "ncclCommInitAll")
        DPQ_SYM(CommDestroy, "ncclCommDestroy")
        DPQ_SYM(AllGather, "ncclAllGather")
        DPQ_SYM(GroupStart, "ncclGroupStart")
        DPQ_SYM(GroupEnd, "ncclGroupEnd")
        DPQ_SYM(GetErrorString, "ncclGetErrorString")
#undef DPQ_SYM
        return std::string();
    }
};
Nccl g_nccl;

}  // namespace

struct dpq_multi {
    int n = 0, M = 0, K = 0, Ds = 0;
    std::vector<dpq_index*> ix;
    std::vector<cudaStream_t> st;
    std::vector<ncclComm_t> comm;
    std::vector<void*> d_q, d_loc, d_all;
    void* d_out = nullptr;  // device 0
    size_t q_cap = 0, k_cap = 0;
    void* h_stage = nullptr;
    size_t h_cap = 0;
    std::vector<uint32_t> pos2id;  // whole tree
};

#define CUM(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return dpq::api_fail(DPQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define NCM(call)                                                                                        \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != 0) return dpq::api_fail(DPQ_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

extern "C" {

int dpq_multi_open_file(const char* tree_path, const char* qnode_path, int M, int K, int n_gpus, dpq_multi** out) {
    if (!tree_path || !out || n_gpus < 1) return dpq::api_fail(DPQ_ERR_ARG, "dpq_multi_open_file: bad argument");
    *out = nullptr;
    if (n_gpus > dpq_device_count())
        return dpq::api_fail(DPQ_ERR_CUDA, "dpq_multi_open_file: fewer CUDA devices than -gpus");
    if (n_gpus > 1) {
        std::string err = g_nccl.load();
        if (!err.empty()) return dpq::api_fail(DPQ_ERR_CUDA, err);
    }
    dpq_multi* m = new dpq_multi();
    m->n = n_gpus;
    m->M = M;
    m->K = K;
    m->ix.assign((size_t)n_gpus, nullptr);
    m->st.assign((size_t)n_gpus, nullptr);
    m->d_q.assign((size_t)n_gpus, nullptr);
    m->d_loc.assign((size_t)n_gpus, nullptr);
    m->d_all.assign((size_t)n_gpus, nullptr);
    for (int r = 0; r < n_gpus; ++r) {
        int rc = dpq_set_device(r);
        if (!rc) rc = dpq_index_open_file(tree_path, nullptr, M, K, r, n_gpus, &m->ix[(size_t)r]);
        if (rc) {
            dpq_multi_close(m);
            return rc;
        }
        cudaSetDevice(r);
        if (cudaStreamCreateWithFlags(&m->st[(size_t)r], cudaStreamNonBlocking) != cudaSuccess ||
            dpq_index_set_stream(m->ix[(size_t)r], m->st[(size_t)r]) != DPQ_OK) {
            dpq_multi_close(m);
            return dpq::api_fail(DPQ_ERR_CUDA, "dpq_multi_open_file: stream setup failed");
        }
    }
    dpq_set_device(0);
    if (qnode_path) {  // vec_id of every position (60-byte QNode records, DCAT.h:79-101)
        FILE* f = fopen(qnode_path, "rb");
        if (!f) {
            dpq_multi_close(m);
            return dpq::api_fail(DPQ_ERR_IO, std::string("cannot open ") + qnode_path);
        }
        const int64_t n_codes = dpq_index_stat(m->ix[0], "n_codes");
        m->pos2id.resize((size_t)n_codes);
        std::vector<uint8_t> rec(60 * 4096);
        size_t done = 0;
        bool ok = true;
        while (ok && done < m->pos2id.size()) {
            size_t want = std::min<size_t>(4096, m->pos2id.size() - done);
            ok = fread(rec.data(), 60, want, f) == want;
            for (size_t i = 0; ok && i < want; ++i) memcpy(&m->pos2id[done + i], rec.data() + 60 * i, 4);
            done += want;
        }
        fclose(f);
        if (!ok) {
            dpq_multi_close(m);
            return dpq::api_fail(DPQ_ERR_FORMAT, "QNode file truncated");
        }
    }
    if (n_gpus > 1) {
        m->comm.assign((size_t)n_gpus, nullptr);
        std::vector<int> devs((size_t)n_gpus);
        for (int r = 0; r < n_gpus; ++r) devs[(size_t)r] = r;
        ncclResult_t r_ = g_nccl.CommInitAll(m->comm.data(), n_gpus, devs.data());
        if (r_ != 0) {
            m->comm.clear();
            dpq_multi_close(m);
            return dpq::api_fail(DPQ_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r_));
        }
    }
    *out = m;
    return DPQ_OK;
}

int dpq_multi_set_codebook(dpq_multi* m, const float* codewords, int Ds) {
    if (!m || !codewords) return dpq::api_fail(DPQ_ERR_ARG, "dpq_multi_set_codebook: null");
    m->Ds = Ds;
    for (int r = 0; r < m->n; ++r) {
        int rc = dpq_index_set_codebook(m->ix[(size_t)r], codewords, Ds);
        if (rc) return rc;
    }
    return DPQ_OK;
}

int dpq_multi_search(dpq_multi* m, const float* queries, int Q, int topk, uint32_t* out_pos, uint32_t* out_id,
                     float* out_dist) {
    if (!m || !queries || Q < 1 || topk < 1) return dpq::api_fail(DPQ_ERR_ARG, "dpq_multi_search: bad argument");
    if (m->Ds < 1) return dpq::api_fail(DPQ_ERR_ARG, "dpq_multi_search: codebook not set");
    const size_t qbytes = (size_t)Q * m->M * m->Ds * 4, kbytes = (size_t)Q * topk * 8;
    if (m->h_cap < qbytes + kbytes) {
        if (m->h_stage) cudaFreeHost(m->h_stage);
        m->h_stage = nullptr;
        CUM(cudaMallocHost(&m->h_stage, qbytes + kbytes));
        m->h_cap = qbytes + kbytes;
    }
    if (m->q_cap < qbytes || m->k_cap < kbytes) {
        for (int r = 0; r < m->n; ++r) {
            CUM(cudaSetDevice(r));
            for (void** p : {&m->d_q[(size_t)r], &m->d_loc[(size_t)r], &m->d_all[(size_t)r]})
                if (*p) {
                    cudaFree(*p);
                    *p = nullptr;
                }
            CUM(cudaMalloc(&m->d_q[(size_t)r], qbytes));
            CUM(cudaMalloc(&m->d_loc[(size_t)r], kbytes));
            CUM(cudaMalloc(&m->d_all[(size_t)r], kbytes * (size_t)m->n));
        }
        CUM(cudaSetDevice(0));
        if (m->d_out) cudaFree(m->d_out);
        CUM(cudaMalloc(&m->d_out, kbytes));
        m->q_cap = qbytes;
        m->k_cap = kbytes;
    }
    float* hq = reinterpret_cast<float*>(m->h_stage);
    uint64_t* hk = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(m->h_stage) + qbytes);
    memcpy(hq, queries, qbytes);
    // every GPU: queries in, local top-k (global positions) of its shard
    for (int r = 0; r < m->n; ++r) {
        CUM(cudaSetDevice(r));
        CUM(cudaMemcpyAsync(m->d_q[(size_t)r], hq, qbytes, cudaMemcpyHostToDevice, m->st[(size_t)r]));
        int rc = dpq_index_search_device(m->ix[(size_t)r], reinterpret_cast<const float*>(m->d_q[(size_t)r]), Q, topk,
                                         reinterpret_cast<uint64_t*>(m->d_loc[(size_t)r]));
        if (rc) return rc;
    }
    const void* merged = m->d_loc[0];
    if (m->n > 1) {  // the only collective: all-gather of the key lists, then the k-way merge
        NCM(g_nccl.GroupStart());
        for (int r = 0; r < m->n; ++r)
            NCM(g_nccl.AllGather(m->d_loc[(size_t)r], m->d_all[(size_t)r], (size_t)Q * topk, kNcclUint64,
                                 m->comm[(size_t)r], m->st[(size_t)r]));
        NCM(g_nccl.GroupEnd());
        CUM(cudaSetDevice(0));
        int rc = dpq_merge_topk_device(m->ix[0], reinterpret_cast<const uint64_t*>(m->d_all[0]), m->n, Q, topk,
                                       reinterpret_cast<uint64_t*>(m->d_out));
        if (rc) return rc;
        merged = m->d_out;
    }
    CUM(cudaSetDevice(0));
    CUM(cudaMemcpyAsync(hk, merged, kbytes, cudaMemcpyDeviceToHost, m->st[0]));
    for (int r = 0; r < m->n; ++r) {
        int rc = dpq_index_sync(m->ix[(size_t)r]);  // stream sync + fallback overflow check
        if (rc) return rc;
    }
    for (size_t i = 0; i < (size_t)Q * topk; ++i) {
        const uint32_t pos = (uint32_t)hk[i];
        const uint32_t bits = (uint32_t)(hk[i] >> 32);
        if (out_pos) out_pos[i] = pos;
        if (out_dist) memcpy(&out_dist[i], &bits, 4);
        if (out_id) out_id[i] = (!m->pos2id.empty() && pos != 0xFFFFFFFFu) ? m->pos2id[pos] : pos;
    }
    return DPQ_OK;
}

int64_t dpq_multi_stat(dpq_multi* m, int rank, const char* name) {
    if (!m || rank < 0 || rank >= m->n) return -1;
    return dpq_index_stat(m->ix[(size_t)rank], name);
}

void dpq_multi_close(dpq_multi* m) {
    if (!m) return;
    for (size_t r = 0; r < m->comm.size(); ++r)
        if (m->comm[r]) g_nccl.CommDestroy(m->comm[r]);
    for (int r = 0; r < m->n; ++r) {
        cudaSetDevice(r);
        if (m->ix[(size_t)r]) dpq_index_close(m->ix[(size_t)r]);
        if (m->st[(size_t)r]) cudaStreamDestroy(m->st[(size_t)r]);
        for (void* p : {m->d_q[(size_t)r], m->d_loc[(size_t)r], m->d_all[(size_t)r]})
            if (p) cudaFree(p);
    }
    cudaSetDevice(0);
    if (m->d_out) cudaFree(m->d_out);
    if (m->h_stage) cudaFreeHost(m->h_stage);
    (void)cudaGetLastError();
    dpq_set_device(0);
    delete m;
}

}  // extern "C"
