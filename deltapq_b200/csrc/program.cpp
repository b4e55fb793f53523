// Host-side tree compiler: DeltaTree byte stream -> device scan program (see
// dpq_internal.h).  The stream format is the reference's (writer
// DCAT.h:1765-1842, reader DCAT.h:3773-3882): int64 header stripped by the caller, M root
// bytes, then per pair of nodes one depth byte (two nibbles) followed by each node's
// changed-subspace bitmap and new centroid bytes; a trailing single node carries a full
// depth byte.  M > 8 uses the extension format (ceil(M/8) bitmap bytes, little-endian;
// 4-bit depths) because the reference's own M = 16 output is invalid.
#include "dpq_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace dpq {
namespace {

struct Emitter {
    ScanProgram* p;
    int M, K, levels;
    OpFormat fmt;
    int chunk_nodes;
    // current chunk
    bool open = false;
    int nodes_in_chunk = 0;
    size_t chunk_op_begin = 0;
    long prev_last = -1;  // index of the previous node's header word (this chunk), or -1
    int prev_depth = 0;
    std::vector<long> level_idx;  // header word index of the in-chunk node at each depth

    void begin(uint32_t first_pos, int depth, const uint8_t* stack, bool emit_root) {
        ChunkDesc c;
        chunk_op_begin = p->ops.size();
        c.quad_begin = (uint32_t)(chunk_op_begin / 4);
        c.n_quads = 0;
        c.first_pos = first_pos;
        c.n_anc_flags = (uint32_t)depth | (emit_root ? CHUNK_EMIT_ROOT : 0u);
        p->chunks.push_back(c);
        size_t a0 = p->anc.size();
        p->anc.resize(a0 + (size_t)levels * M, 0);
        memcpy(p->anc.data() + a0, stack, (size_t)depth * M);
        level_idx.assign((size_t)levels + 1, -1);
        prev_last = -1;
        nodes_in_chunk = 0;
        open = true;
    }
    void end() {
        if (!open) return;
        p->chunks.back().n_quads = (uint32_t)((p->ops.size() - chunk_op_begin) / 4);
        open = false;
    }
    // node at `depth` whose parent code is `par`, own code `cur`
    void node(int depth, const uint8_t* par, const uint8_t* cur) {
        if (prev_last >= 0) {  // now the previous node's successor is known: patch its header
            uint32_t& w = p->ops[(size_t)prev_last];
            if (depth == prev_depth + 1) {
                w |= OP_CHILD;
            } else if (depth < prev_depth) {
                int lev = depth - 1;
                w |= OP_AUX | fmt.level_bits((uint32_t)lev);
                long owner = level_idx[(size_t)lev];
                if (owner >= 0)  // the ancestor at `lev` lives in this chunk: make it store
                    p->ops[(size_t)owner] |= OP_AUX | fmt.level_bits((uint32_t)lev);
            }
        }
        const size_t first = p->ops.size();
        int n = 0;
        for (int m = 0; m < M; ++m)
            if (par[m] != cur[m]) {
                uint32_t fr = (uint32_t)(m * K + par[m]) << 2;
                uint32_t to = (uint32_t)(m * K + cur[m]) << 2;
                p->ops.push_back(fr | (to << fmt.tshift()));
                ++n;
            }
        const int nq = n == 0 ? 1 : (n + 3) / 4;
        p->ops.resize(first + (size_t)nq * 4, 0u);  // zero ops: table row 0 minus row 0
        p->ops[first] |= (uint32_t)(nq - 1) << 30;
        prev_last = (long)first;  // header lives in the record's first word
        prev_depth = depth;
        level_idx[(size_t)depth] = prev_last;
        ++nodes_in_chunk;
    }
};

}  // namespace

std::string compile_program(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M, int K,
                            int rank, int n_ranks, int chunk_nodes, ScanProgram* out, int engine) {
    if (M < 1 || M > 16 || K < 1 || K > 256) return "unsupported M/K (need 1<=M<=16, 1<=K<=256)";
    if (n_codes < 1) return "empty tree";
    if (n_codes >= 0x7FFFFFFFLL) return "n_codes must be < 2^31-1 (DCAT.h:982)";
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return "bad rank / n_ranks";
    if (n_bytes < M) return "stream shorter than the root code";
    ScanProgram& P = *out;
    P = ScanProgram();
    P.M = M;
    P.K = K;
    P.fmt.rb = (M * K <= 2048) ? 11 : 12;
    if (M * K > 4096) return "M*K > 4096 not supported";
    P.n_codes = n_codes;
    P.n_bytes = n_bytes;
    P.v2 = engine == 0 && v2_shape_ok(M, K);
    if (P.v2) P.shape = v2_shape(M, K);
    // Depth limit of the FORMAT: a nibble, masked with &7 by the reference reader when M <= 8
    // (DCAT.h:3794), 4 bits in the M > 8 extension.  The first-generation engine additionally keeps a
    // per-warp depth stack of fmt.levels() entries; the code-array engine (v2) has no stack.
    const int bmb = (M + 7) / 8;
    const int dmask = M > 8 ? 15 : 7;
    const int levels = P.v2 ? (M > 8 ? 16 : 8) : P.fmt.levels();
    P.depth_hist.assign((size_t)std::max(levels, P.fmt.levels()) + 1, 0);
    if (chunk_nodes < 4) chunk_nodes = 4;
    const int cs = P.v2 ? P.shape.nf : M;
    P.cstride = cs;
    uint8_t padded[16] = {0};

    // capacity up front (a shard holds about 1 / n_ranks of the nodes): no regrowth copies of GB-sized arrays
    {
        const size_t share = (size_t)(n_codes / n_ranks + n_codes / (8 * n_ranks) + 1024);
        const size_t cap = std::min<size_t>((size_t)n_codes, share);
        P.codes.reserve(cap * (size_t)cs);
    }
    std::vector<uint8_t> stack((size_t)(levels + 1) * M, 0);
    int64_t off = 0;
    memcpy(stack.data(), payload, (size_t)M);
    off = M;

    Emitter E;
    E.p = &P;
    E.M = M;
    E.K = K;
    E.levels = levels;
    E.fmt = P.fmt;
    E.chunk_nodes = chunk_nodes;

    bool have_base = false;
    bool pending_root = (rank == 0);
    if (rank == 0) {
        P.base_pos = 0;
        have_base = true;
        memcpy(padded, payload, (size_t)M);
        P.codes.insert(P.codes.end(), padded, padded + cs);  // the root is an ordinary node at position 0
        P.n_local = 1;
        P.local_bytes = M;
        P.depth_hist[0] = 1;
    }
    int cur_rank = 0;
    int depths = 0;
    int last_depth = 0;
    int64_t local_nodes_records = 0;
    for (int64_t i = 1; i < n_codes; ++i) {
        int64_t rec_begin = off;
        int d;
        if (i & 1) {
            if (off >= n_bytes) return "stream truncated (depth byte)";
            depths = payload[off++];
            d = (i == n_codes - 1) ? depths : (depths & dmask);  // DCAT.h:3861 unmasked tail
        } else {
            d = (depths >> 4) & dmask;
        }
        if (d < 1 || d > levels - 1 || d > last_depth + 1) return "bad depth in stream";
        last_depth = d;
        if (off + bmb > n_bytes) return "stream truncated (bitmap)";
        uint32_t bitmap = 0;
        for (int b = 0; b < bmb; ++b) bitmap |= (uint32_t)payload[off++] << (8 * b);
        if (M < 32 && (bitmap >> M)) return "bitmap has bits above M";
        uint8_t* cur = stack.data() + (size_t)d * M;
        const uint8_t* par = stack.data() + (size_t)(d - 1) * M;
        memcpy(cur, par, (size_t)M);
        int nd = 0;
        for (int m = 0; m < M; ++m)
            if ((bitmap >> m) & 1) {
                if (off >= n_bytes) return "stream truncated (centroid byte)";
                uint8_t c = payload[off++];
                if (c >= K) return "centroid id >= K in stream";
                cur[m] = c;
                ++nd;
            }
        if (d == 1) {  // a depth-1 subtree starts: pick its owner by stream byte offset
            cur_rank = (int)((__int128)rec_begin * n_ranks / n_bytes);
            if (cur_rank >= n_ranks) cur_rank = n_ranks - 1;
        }
        if (cur_rank != rank) {
            if (E.open) E.end();
            continue;
        }
        if (!have_base) {
            P.base_pos = i;
            have_base = true;
        }
        if (P.v2) {
            // a shard is one contiguous range of positions: whole depth-1 subtrees are dealt out in
            // stream order, and rank 0 holds the root (position 0) in front of its subtrees
            if (P.base_pos + P.n_local != i) return "internal: shard positions not contiguous";
            memcpy(padded, cur, (size_t)M);
            P.codes.insert(P.codes.end(), padded, padded + cs);
        } else {
            if (!E.open || E.nodes_in_chunk >= chunk_nodes) {
                E.end();
                E.begin((uint32_t)i, d, stack.data(), pending_root);
                pending_root = false;
            }
            E.node(d, par, cur);
            P.codes.insert(P.codes.end(), cur, cur + M);
        }
        P.n_local++;
        P.n_diffs += nd;
        P.local_bytes += bmb + nd;
        local_nodes_records++;
        P.depth_hist[(size_t)d]++;
    }
    E.end();
    if (off != n_bytes) return "stream has trailing or missing bytes (n_bytes mismatch)";
    P.local_bytes += (local_nodes_records + 1) / 2;  // depth nibbles
    if (pending_root && !P.v2) {  // root only (n_codes == 1, or rank 0 owns no subtree)
        E.begin(1u, 1, stack.data(), true);
        E.end();
    }
    if (!have_base) P.base_pos = n_codes;
    return std::string();
}

}  // namespace dpq
