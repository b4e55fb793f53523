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

// v2 emitter: one fixed record per node (see dpq_internal.h)
struct Emitter2 {
    ScanProgram* p;
    int M, K;
    V2Shape sh;
    int chunk_nodes;
    bool open = false;
    int in_chunk = 0;
    long prev_rec = -1;     // word index of the previous record in this chunk, or -1
    int prev_depth = 0;
    long reg_owner_depth = -1;  // depth of the node whose distance the parent register holds, -1: none
    // the register holds dist(node at depth reg_owner_depth on the current path) iff that
    // node is still on the path, i.e. no shallower-or-equal node was emitted since
    void begin(uint32_t first_pos) {
        ChunkDesc2 c;
        c.rec_begin = (uint32_t)(p->recs.size() / (size_t)sh.rec_words());
        c.n_nodes = 0;
        c.first_pos = first_pos;
        c.pad = 0;
        p->chunks2.push_back(c);
        open = true;
        in_chunk = 0;
        prev_rec = -1;
        reg_owner_depth = -1;
    }
    void end() {
        if (!open) return;
        p->chunks2.back().n_nodes = (uint32_t)in_chunk;
        open = false;
    }
    void node(int depth, const uint8_t* par, const uint8_t* cur) {
        if (prev_rec >= 0 && depth == prev_depth + 1) {  // previous node is this node's parent
            p->recs[(size_t)prev_rec] |= V2_CHILD;
            reg_owner_depth = prev_depth;
        } else if (reg_owner_depth >= depth) {
            reg_owner_depth = -1;  // the register's node left the path
        }
        const int nf = sh.nf, half = sh.nf / 2;
        int nd = 0;
        for (int m = 0; m < M; ++m) nd += par[m] != cur[m];
        uint32_t f[16];
        const bool abs = !(reg_owner_depth >= 0 && reg_owner_depth == depth - 1 && nd <= half);
        if (!abs) {
            for (int i = 0; i < nf; ++i) f[i] = 0;
            int j = 0;
            for (int m = 0; m < M; ++m)
                if (par[m] != cur[m]) {
                    f[j] = (uint32_t)(m * K + cur[m]) * (uint32_t)sh.lpg;         // plus: new centroid
                    f[half + j] = (uint32_t)(m * K + par[m]) * (uint32_t)sh.lpg;  // minus: old centroid
                    ++j;
                }
            p->v2_delta_nodes++;
        } else {
            const uint32_t zero_row = (uint32_t)(M * K) * (uint32_t)sh.lpg;
            for (int i = 0; i < nf; ++i) f[i] = i < M ? (uint32_t)(i * K + cur[i]) * (uint32_t)sh.lpg : zero_row;
        }
        const size_t at = p->recs.size();
        p->recs.resize(at + (size_t)half);
        uint32_t* r = p->recs.data() + at;
        for (int w = 0; w < half; ++w) r[w] = f[2 * w] | (f[2 * w + 1] << 16);
        if (abs) r[0] |= V2_ABS;
        prev_rec = (long)at;
        prev_depth = depth;
        ++in_chunk;
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
    const int levels = P.fmt.levels();
    const int bmb = (M + 7) / 8;
    const int dmask = M > 8 ? 15 : 7;  // DCAT.h:3794 masks nibbles with &7
    P.depth_hist.assign((size_t)levels + 1, 0);
    if (chunk_nodes < 4) chunk_nodes = 4;
    P.v2 = engine == 0 && v2_shape_ok(M, K);
    if (P.v2) P.shape = v2_shape(M, K);
    Emitter2 E2;
    E2.p = &P;
    E2.M = M;
    E2.K = K;
    E2.sh = P.shape;
    E2.chunk_nodes = P.v2_chunk_nodes;

    // capacity up front (a shard holds about 1 / n_ranks of the nodes): no regrowth copies of GB-sized arrays
    {
        const size_t share = (size_t)(n_codes / n_ranks + n_codes / (8 * n_ranks) + 1024);
        const size_t cap = std::min<size_t>((size_t)n_codes, share);
        P.codes.reserve(cap * (size_t)M);
        if (P.v2) {
            P.recs.reserve(cap * (size_t)P.shape.rec_words());
            P.chunks2.reserve(cap / (size_t)P.v2_chunk_nodes + 16);
        }
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
        P.codes.insert(P.codes.end(), payload, payload + M);
        P.n_local = 1;
        P.local_bytes = M;
        P.depth_hist[0] = 1;
        if (P.v2) {  // the root is an ordinary full record at position 0
            E2.begin(0u);
            E2.node(0, payload, payload);
        }
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
            if (E2.open) E2.end();
            continue;
        }
        if (!have_base) {
            P.base_pos = i;
            have_base = true;
        }
        if (P.v2) {
            // positions of one shard are contiguous except that rank 0 also holds the root
            const bool gap = E2.open && P.chunks2.back().first_pos + (uint32_t)E2.in_chunk != (uint32_t)i;
            if (!E2.open || E2.in_chunk >= E2.chunk_nodes || gap) {
                E2.end();
                E2.begin((uint32_t)i);
            }
            E2.node(d, par, cur);
        } else {
            if (!E.open || E.nodes_in_chunk >= chunk_nodes) {
                E.end();
                E.begin((uint32_t)i, d, stack.data(), pending_root);
                pending_root = false;
            }
            E.node(d, par, cur);
        }
        P.codes.insert(P.codes.end(), cur, cur + M);
        P.n_local++;
        P.n_diffs += nd;
        P.local_bytes += bmb + nd;
        local_nodes_records++;
        P.depth_hist[(size_t)d]++;
    }
    E.end();
    E2.end();
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
