// Device-side tree compiler of the code-array engine (dpq_internal.h "v2").
//
//  * launch_pad_codes: codes by position [n][M] -> the scan's word stride (8 or 16 bytes per node).
//  * decode_stream_device: the on-disk DeltaTree stream (writer DCAT.h:1765-1842,
//    reader DCAT.h:3773-3882) decoded ON THE GPU into a device-resident dpq_tree (codes by position,
//    depth, record offsets).  The stream is a strictly sequential byte code -- variable-length
//    records, a node's parent is the most recent node one level up -- which one host core decodes at
//    ~14M nodes/s (8.9 s per 125M nodes, program.cpp).  The parallel form:
//      1. record boundaries: the stream after the root is cut into 4 KB blocks; for every block and
//         every possible entry offset (a pair of records is at most 1 + 2 (bitmap + M) bytes long, so
//         a pair boundary falls within that many bytes of a block start) a thread walks the block's
//         pairs and records (exit offset into the next block, pairs walked).  The host then chains
//         the blocks' transition tables (one table lookup per block) to the true entry offset and
//         first pair index of every block;
//      2. every block re-walks its pairs from its true entry and writes per node: depth, changed-
//         subspace bitmap, stream offset of its record;
//      3. parent of node i = the last node before i at depth(i) - 1: one inclusive max-scan per tree
//         level (cub) of "position if the node is at this level";
//      4. codes level by level: a node copies its parent's code (final, one level up) and overwrites
//         the changed subspaces with its record's bytes.
//    Malformed streams (truncation, bad depth, centroid id >= K, bitmap bits above M, length
//    mismatch) are detected on the device and reported like the host decoder's errors.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "../../include/dpq.h"
#include "dpq_internal.h"
#include "kernels.cuh"
#include "tree_internal.h"

namespace dpq {

int api_fail(int code, const std::string& msg);
int api_check_device();
int api_device();

// codes [n][M] -> out [n][stride] (stride 8 or 16), pad bytes 0; one thread per output 32-bit word
__global__ void pad_codes_kernel(const uint8_t* __restrict__ codes, int64_t n, int M, int stride,
                                 uint32_t* __restrict__ out) {
    const int wpn = stride / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * wpn) return;
    const int64_t node = i / wpn;
    const int b0 = (int)(i % wpn) * 4;
    const uint8_t* c = codes + (size_t)node * M;
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (b0 + b < M) w |= (uint32_t)c[b0 + b] << (8 * b);
    out[i] = w;
}

cudaError_t launch_pad_codes(const uint8_t* codes, int64_t n, int M, int stride, uint8_t* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t words = n * (stride / 4);
    const unsigned blocks = (unsigned)((words + 255) / 256);
    pad_codes_kernel<<<blocks, 256, 0, st>>>(codes, n, M, stride, reinterpret_cast<uint32_t*>(out));
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
namespace {

constexpr int DEC_BLOCK = 4096;   // stream bytes per block
constexpr uint32_t E_TRUNC = 1u, E_DEPTH = 2u, E_CENTROID = 4u, E_BITMAP = 8u, E_LENGTH = 16u;
constexpr uint16_t EXIT_NONE = 0xFFFFu;

#define CUD(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return api_fail(DPQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct DBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
    void* take() {
        void* q = p;
        p = nullptr;
        return q;
    }
    ~DBuf() {
        if (p) cudaFree(p);
    }
};

__device__ __forceinline__ uint32_t read_bitmap(const uint8_t* __restrict__ s, int64_t o, int bmb) {
    uint32_t bm = s[o];
    if (bmb == 2) bm |= (uint32_t)s[o + 1] << 8;
    return bm;
}

// step 1: block b, candidate entry e (pair boundary at block_start + e): walk pairs to the block end
__global__ void dec_spec_kernel(const uint8_t* __restrict__ s, int64_t n_bytes, int M, int bmb, int W, int64_t n_blocks,
                                uint16_t* __restrict__ t_exit, uint16_t* __restrict__ t_cnt) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_blocks * W) return;
    const int64_t b = tid / W;
    const int e = (int)(tid % W);
    const int64_t begin = (int64_t)M + b * DEC_BLOCK, end = begin + DEC_BLOCK;
    int64_t o = begin + e;
    uint32_t cnt = 0;
    bool over = false;
    while (o < end && o < n_bytes) {
        int64_t p = o + 1;  // the pair's depth byte
        for (int k = 0; k < 2; ++k) {
            if (p + bmb > n_bytes) {
                over = true;
                break;
            }
            p += bmb + __popc(read_bitmap(s, p, bmb));
        }
        if (over) break;
        o = p;
        ++cnt;
    }
    t_exit[tid] = (over || o < end) ? EXIT_NONE : (uint16_t)(o - end);  // o < end: the stream ended inside this block
    t_cnt[tid] = (uint16_t)cnt;
}

// step 2: block b walks its pairs from its true entry; pair j holds nodes 2j+1 and 2j+2
__global__ void dec_emit_kernel(const uint8_t* __restrict__ s, int64_t n_bytes, int64_t n, int M, int bmb, int dmask,
                                int64_t n_blocks, const int32_t* __restrict__ entry, const int64_t* __restrict__ first_pair,
                                uint8_t* __restrict__ depth, uint16_t* __restrict__ bitmap,
                                unsigned long long* __restrict__ roff, uint32_t* __restrict__ err) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks || entry[b] < 0) return;
    const int64_t begin = (int64_t)M + b * DEC_BLOCK, end = begin + DEC_BLOCK;
    int64_t o = begin + entry[b];
    int64_t j = first_pair[b];
    while (o < end) {
        const int64_t i1 = 2 * j + 1;
        if (i1 >= n) {  // all nodes decoded: nothing may follow
            if (o != n_bytes) atomicOr(err, E_LENGTH);
            return;
        }
        if (o >= n_bytes) {
            atomicOr(err, E_TRUNC);
            return;
        }
        const uint32_t depths = s[o];
        int64_t p = o + 1;
        for (int k = 0; k < 2; ++k) {
            const int64_t i = i1 + k;
            if (i >= n) break;
            if (p + bmb > n_bytes) {
                atomicOr(err, E_TRUNC);
                return;
            }
            const uint32_t bm = read_bitmap(s, p, bmb);
            if (bm >> M) atomicOr(err, E_BITMAP);
            // DCAT.h:3794 masks the nibbles with &7 (M <= 8); the trailing single node's byte is unmasked (:3861)
            const uint32_t d = k == 0 ? ((i == n - 1) ? depths : (depths & (uint32_t)dmask)) : ((depths >> 4) & (uint32_t)dmask);
            depth[i] = (uint8_t)min(d, 255u);
            bitmap[i] = (uint16_t)bm;
            roff[i - 1] = (unsigned long long)((k == 0 ? o : p) - M);  // record start: the depth byte for odd nodes
            p += bmb + __popc(bm);
            if (p > n_bytes) {
                atomicOr(err, E_TRUNC);
                return;
            }
            if (i == n - 1 && p != n_bytes) atomicOr(err, E_LENGTH);
        }
        o = p;
        ++j;
    }
}

// depth rules (program.cpp): 1 <= d <= max_level, d <= previous depth + 1; 0xFF = never written
__global__ void dec_depth_check_kernel(const uint8_t* __restrict__ depth, int64_t n, int max_level, uint32_t* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= n) return;
    const int d = depth[i], prev = depth[i - 1];
    if (d == 0xFF) atomicOr(err, E_LENGTH);
    else if (d < 1 || d > max_level || d > prev + 1) atomicOr(err, E_DEPTH);
}

struct LevelPos {  // position of the node if it is at `level`, else 0 (the root is every level-1 node's fallback)
    const uint8_t* depth;
    int level;
    __device__ uint32_t operator()(uint32_t i) const { return depth[i] == level ? i : 0u; }
};

// nodes at level + 1 take the last node at `level` seen strictly before them
__global__ void dec_parent_kernel(const uint8_t* __restrict__ depth, const uint32_t* __restrict__ last_at, int64_t n, int level,
                                  uint32_t* __restrict__ parent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= n || depth[i] != level + 1) return;
    parent[i] = last_at[i - 1];
}

// level by level: parent's code with the changed subspaces replaced by the record's bytes
__global__ void dec_codes_kernel(const uint8_t* __restrict__ s, const uint8_t* __restrict__ depth,
                                 const uint16_t* __restrict__ bitmap, const unsigned long long* __restrict__ roff,
                                 const uint32_t* __restrict__ parent, int64_t n, int M, int K, int bmb, int level,
                                 uint8_t* __restrict__ codes, uint32_t* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= n || depth[i] != level) return;
    const uint8_t* pc = codes + (size_t)parent[i] * M;
    uint8_t* c = codes + (size_t)i * M;
    const uint32_t bm = bitmap[i];
    int64_t p = (int64_t)roff[i - 1] + M + (i & 1) + bmb;  // first changed byte
    for (int m = 0; m < M; ++m) {
        uint8_t v = pc[m];
        if ((bm >> m) & 1u) {
            v = s[p++];
            if (v >= K) atomicOr(err, E_CENTROID);
        }
        c[m] = v;
    }
}

__global__ void iota32_kernel(uint32_t* a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}

inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

// payload: HOST pointer to the stream without its 16-byte header; pos2id: host, nullable.
// *out: a device-resident dpq_tree (codes by position, depth, record offsets, the stream itself).
int decode_stream_device(const uint8_t* payload, int64_t n_bytes, int64_t n, int M, int K, const uint32_t* pos2id,
                         dpq_tree** out) {
    *out = nullptr;
    if (M < 1 || M > 16 || K < 1 || K > 256) return api_fail(DPQ_ERR_FORMAT, "unsupported M/K (need 1<=M<=16, 1<=K<=256)");
    if (n < 1) return api_fail(DPQ_ERR_FORMAT, "empty tree");
    if (n >= 0x7FFFFFFFLL) return api_fail(DPQ_ERR_FORMAT, "n_codes must be < 2^31-1 (DCAT.h:982)");
    if (n_bytes < M) return api_fail(DPQ_ERR_FORMAT, "stream shorter than the root code");
    int rc = api_check_device();
    if (rc) return rc;
    CUD(cudaSetDevice(api_device()));
    const int bmb = (M + 7) / 8;
    const int dmask = M > 8 ? 15 : 7;
    const int max_level = M > 8 ? 15 : 7;
    const int W = 1 + 2 * (bmb + M);  // longest pair of records
    const int64_t body = n_bytes - M;
    const int64_t n_blocks = std::max<int64_t>((body + DEC_BLOCK - 1) / DEC_BLOCK, 1);

    DBuf d_s, d_depth, d_bitmap, d_roff, d_err, d_codes, d_vec;
    CUD(d_s.alloc((size_t)n_bytes));
    CUD(cudaMemcpy(d_s.p, payload, (size_t)n_bytes, cudaMemcpyHostToDevice));
    CUD(d_depth.alloc((size_t)n));
    CUD(d_bitmap.alloc((size_t)n * 2));
    CUD(d_roff.alloc((size_t)std::max<int64_t>(n - 1, 1) * 8));
    CUD(d_err.alloc(16));
    CUD(cudaMemset(d_err.p, 0, 16));
    CUD(cudaMemset(d_depth.p, 0xFF, (size_t)n));
    CUD(cudaMemset(d_depth.p, 0, 1));
    uint32_t* err = d_err.as<uint32_t>();
    const uint8_t* s = d_s.as<uint8_t>();

    if (n > 1) {
        if (body < 1) return api_fail(DPQ_ERR_FORMAT, "stream truncated (depth byte)");
        {   // steps 1 + 2: record boundaries
            DBuf d_exit, d_cnt, d_entry, d_first;
            CUD(d_exit.alloc((size_t)n_blocks * W * 2));
            CUD(d_cnt.alloc((size_t)n_blocks * W * 2));
            dec_spec_kernel<<<nblk(n_blocks * W), 256>>>(s, n_bytes, M, bmb, W, n_blocks, d_exit.as<uint16_t>(), d_cnt.as<uint16_t>());
            std::vector<uint16_t> t_exit((size_t)n_blocks * W), t_cnt((size_t)n_blocks * W);
            CUD(cudaMemcpy(t_exit.data(), d_exit.p, t_exit.size() * 2, cudaMemcpyDeviceToHost));
            CUD(cudaMemcpy(t_cnt.data(), d_cnt.p, t_cnt.size() * 2, cudaMemcpyDeviceToHost));
            std::vector<int32_t> entry((size_t)n_blocks, -1);
            std::vector<int64_t> first((size_t)n_blocks, 0);
            int e = 0;
            int64_t pairs = 0;
            for (int64_t b = 0; b < n_blocks && e >= 0; ++b) {  // one table lookup per block
                entry[(size_t)b] = e;
                first[(size_t)b] = pairs;
                const size_t at = (size_t)b * W + (size_t)e;
                pairs += t_cnt[at];
                e = t_exit[at] == EXIT_NONE ? -1 : (int)t_exit[at];
                if (e >= W) e = -1;
            }
            CUD(d_entry.alloc((size_t)n_blocks * 4));
            CUD(d_first.alloc((size_t)n_blocks * 8));
            CUD(cudaMemcpy(d_entry.p, entry.data(), (size_t)n_blocks * 4, cudaMemcpyHostToDevice));
            CUD(cudaMemcpy(d_first.p, first.data(), (size_t)n_blocks * 8, cudaMemcpyHostToDevice));
            dec_emit_kernel<<<nblk(n_blocks), 256>>>(s, n_bytes, n, M, bmb, dmask, n_blocks, d_entry.as<int32_t>(),
                                                    d_first.as<int64_t>(), d_depth.as<uint8_t>(), d_bitmap.as<uint16_t>(),
                                                    d_roff.as<unsigned long long>(), err);
            dec_depth_check_kernel<<<nblk(n - 1), 256>>>(d_depth.as<uint8_t>(), n, max_level, err);
            CUD(cudaGetLastError());
        }
        uint32_t flags = 0;
        CUD(cudaMemcpy(&flags, err, 4, cudaMemcpyDeviceToHost));
        if (flags & E_TRUNC) return api_fail(DPQ_ERR_FORMAT, "stream truncated");
        if (flags & E_LENGTH) return api_fail(DPQ_ERR_FORMAT, "stream has trailing or missing bytes (n_bytes mismatch)");
        if (flags & E_BITMAP) return api_fail(DPQ_ERR_FORMAT, "bitmap has bits above M");
        if (flags & E_DEPTH) return api_fail(DPQ_ERR_FORMAT, "bad depth in stream");
    } else if (n_bytes != M) {
        return api_fail(DPQ_ERR_FORMAT, "stream has trailing or missing bytes (n_bytes mismatch)");
    }

    CUD(d_codes.alloc((size_t)n * M));
    CUD(cudaMemcpy(d_codes.p, s, (size_t)M, cudaMemcpyDeviceToDevice));  // the root's code
    if (n > 1) {
        DBuf d_parent, d_last, d_tmp;
        CUD(d_parent.alloc((size_t)n * 4));
        CUD(d_last.alloc((size_t)n * 4));
        CUD(cudaMemset(d_parent.p, 0, (size_t)n * 4));
        int64_t hist[17];
        if ((rc = depth_hist_device(api_device(), d_depth.as<uint8_t>(), n, hist))) return rc;
        int deepest = 0;
        for (int d = 0; d <= 16; ++d)
            if (hist[d]) deepest = d;
        cub::CountingInputIterator<uint32_t> count(0u);
        size_t tmp_bytes = 0;
        {
            cub::TransformInputIterator<uint32_t, LevelPos, cub::CountingInputIterator<uint32_t>> it(count, LevelPos{d_depth.as<uint8_t>(), 0});
            CUD(cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, it, d_last.as<uint32_t>(), cub::Max(), (int)n));
        }
        CUD(d_tmp.alloc(tmp_bytes));
        for (int level = 0; level < deepest; ++level) {
            if (level > 0) {  // level 0: every depth-1 node hangs off the root (parent stays 0)
                cub::TransformInputIterator<uint32_t, LevelPos, cub::CountingInputIterator<uint32_t>> it(
                    count, LevelPos{d_depth.as<uint8_t>(), level});
                size_t tb = tmp_bytes;
                CUD(cub::DeviceScan::InclusiveScan(d_tmp.p, tb, it, d_last.as<uint32_t>(), cub::Max(), (int)n));
                dec_parent_kernel<<<nblk(n - 1), 256>>>(d_depth.as<uint8_t>(), d_last.as<uint32_t>(), n, level, d_parent.as<uint32_t>());
            }
            dec_codes_kernel<<<nblk(n - 1), 256>>>(s, d_depth.as<uint8_t>(), d_bitmap.as<uint16_t>(), d_roff.as<unsigned long long>(),
                                                  d_parent.as<uint32_t>(), n, M, K, bmb, level + 1, d_codes.as<uint8_t>(), err);
        }
        CUD(cudaGetLastError());
        uint32_t flags = 0;
        CUD(cudaMemcpy(&flags, err, 4, cudaMemcpyDeviceToHost));
        if (flags & E_CENTROID) return api_fail(DPQ_ERR_FORMAT, "centroid id >= K in stream");
    }
    if (K < 256) {  // the root's bytes
        std::vector<uint8_t> root(payload, payload + M);
        for (int m = 0; m < M; ++m)
            if (root[(size_t)m] >= K) return api_fail(DPQ_ERR_FORMAT, "centroid id >= K in stream");
    }
    if (pos2id) {
        CUD(d_vec.alloc((size_t)n * 4));
        CUD(cudaMemcpy(d_vec.p, pos2id, (size_t)n * 4, cudaMemcpyHostToDevice));
    }
    CUD(cudaDeviceSynchronize());
    dpq_tree* t = new dpq_tree();
    t->M = M;
    t->K = K;
    t->n = n;
    t->on_device = true;
    t->device = api_device();
    t->payload_bytes = n_bytes;
    t->n_diffs = n_bytes - M - (int64_t)bmb * (n - 1) - n / 2;  // M = 8: n_bytes = 8 + n_diffs + (3(n-1)+1)/2
    t->d_codes_by_pos = d_codes.take();
    t->d_depth = d_depth.take();
    t->d_vec_id = d_vec.take();  // null without pos2id
    t->d_payload = d_s.take();
    t->d_roff = d_roff.take();
    *out = t;
    return DPQ_OK;
}

}  // namespace dpq
