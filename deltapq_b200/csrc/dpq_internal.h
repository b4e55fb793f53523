// Internal declarations shared by the host-side tree compiler and the CUDA kernels.
// Nothing here crosses the C ABI (include/dpq.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dpq {

// ---------------------------------------------------------------------------------------
// Device layout of one DeltaTree shard: the "scan program".
//
// The on-disk stream (reference writer DCAT.h:1765-1842) is a strictly
// sequential byte code: variable-length records, each node's parent found through a
// depth stack.  At open time the host decodes it once and emits, per chunk of
// `chunk_nodes` consecutive DFS positions, a flat array of node RECORDS that a warp can
// execute with no further decoding.  A record is 1..4 quads (uint4) of 32-bit delta ops,
// one op per subspace the node's edge changes, unused slots zero (a zero op reads table
// row 0 twice and cancels):
//
//   op = from_row4 | to_row4 << (RB+2)
//     from_row4 / to_row4 : (m*K + centroid) * 4, byte offset of the ADC-table row of the
//                           subspace this edge changes: old centroid / new centroid
//   the FIRST word of a record also carries, in bits no op field uses:
//     NQ-1  (bits 30..31) : quads in this record
//     CHILD (bit 29)      : the NEXT node is this node's child (its parent distance = ours)
//     AUX   (bit 28)      : with CHILD: also store our distance to the per-warp depth stack
//                           at LEVEL; without CHILD: the next node's parent is stack[LEVEL]
//                           (neither bit: the next node is a sibling, same parent)
//     LEVEL               : bits 0..1 and RB+2..RB+3 (4 bits)
//
// Every chunk also records the codes of the ancestors of its first node, so a warp starts a
// chunk with full M-term table sums for those ancestors ("full M-term lookup at each root")
// and needs nothing from other chunks.
// ---------------------------------------------------------------------------------------
constexpr uint32_t OP_CHILD = 1u << 29;
constexpr uint32_t OP_AUX = 1u << 28;  // with CHILD: store; without: pop

struct OpFormat {
    int rb;  // row bits: 11 (M*K <= 2048) or 12 (M*K <= 4096)
    uint32_t fmask() const { return ((1u << rb) - 1u) << 2; }
    int tshift() const { return rb + 2; }
    int levels() const { return rb == 11 ? 8 : 16; }
    uint32_t level_bits(uint32_t lev) const { return (lev & 3u) | ((lev >> 2) << (rb + 2)); }
};

struct ChunkDesc {
    uint32_t quad_begin;   // first uint4 of this chunk in the op array
    uint32_t n_quads;
    uint32_t first_pos;    // global DFS position of the first node executed by the ops
    uint32_t n_anc_flags;  // bits 0..7: number of ancestor levels; bit 8: also emit the root
};
constexpr uint32_t CHUNK_EMIT_ROOT = 1u << 8;

// ---------------------------------------------------------------------------------------
// Second-generation layout ("v2", used whenever M <= 16): the device copy of a shard is just
// the nodes' PQ CODES by DFS position, padded to a power-of-two stride (8 bytes for M <= 8,
// 16 bytes for M <= 16, pad bytes 0): 8 B/node at M = 8 -- what plain PQ codes take, against
// 6.7 B/node for the delta-coded stream on disk -- and the SAME array serves
// the scan kernels, the exact re-score and the exact fallback.  Every node gets the full
// M-term lookup: with n changed subspaces the reference's delta rule costs 2n table reads per
// (node, query) (subtract old, add new), and the trees here have n = 4.3 .. 5.2 on average, so
// the delta form would read MORE table rows than the M = 8 of a full lookup (measured in round 1:
// only 1.7 % of the nodes of the 1M-code tree could use a <= 4-pair delta record).  The delta
// rule itself lives on in the first-generation engine below (DPQ_ENGINE=1), which every parity
// test also runs.
//
// Scan tables are laid out [row][queries] with row = m * 256 + centroid (rows of subspaces
// m >= M and of centroids >= K are all zero, so a pad byte reads a zero row):
//   narrow (M <= 8):  2048 rows; 15-bit scan: 7 x 16-byte lanes = 56 queries per CTA, 8-bit
//                     coarse scan: 112 queries per CTA; a strand is a quarter warp (8 lanes,
//                     7 active) walking one 64-node chunk: 4 nodes per warp step;
//   wide   (M <= 16): 4096 rows; 15-bit scan: 3 x 16-byte lanes = 24 queries per CTA (a strand
//                     is 4 lanes, 3 active: 8 nodes per warp step); coarse scan: 4-bit entries,
//                     56-byte rows = 112 queries per CTA (scan8.cu).
struct V2Shape {
    int nf;         // table fields per node (8 or 16) = code stride in bytes
    int lpg;        // active 16-byte lanes per strand of the 15-bit scan = table row bytes / 16
    int sw;         // lanes per strand (8 or 4)
    int rows;       // table rows (nf * 256)
    int qb() const { return lpg * 8; }            // queries per CTA (15-bit scan)
    int row_bytes() const { return lpg * 16; }
    int lut_bytes() const { return rows * row_bytes(); }
    int spw() const { return 32 / sw; }           // strands (= chunks in flight) per warp
};
inline V2Shape v2_shape(int M, int K) {
    (void)K;
    if (M <= 8) return V2Shape{8, 7, 8, 2048};
    return V2Shape{16, 3, 4, 4096};
}

struct ScanProgram {
    int M = 0, K = 0;
    bool v2 = false;
    int v2_chunk_nodes = 64;
    V2Shape shape{8, 7, 8, 2048};
    int cstride = 0;                   // bytes per node in `codes` (v2: 8 or 16, zero padded; else M)
    OpFormat fmt{11};
    int64_t n_codes = 0;       // nodes in the whole tree
    int64_t n_bytes = 0;       // stream bytes of the whole tree
    int64_t base_pos = 0;      // first global position held by this shard
    int64_t n_local = 0;       // nodes held by this shard (contiguous positions)
    int64_t local_bytes = 0;   // algorithmic stream bytes of this shard
    int64_t n_diffs = 0;       // changed subspaces over this shard's nodes
    std::vector<uint32_t> ops;        // node records, each a multiple of 4 words
    std::vector<ChunkDesc> chunks;
    std::vector<uint8_t> anc;         // [n_chunks][levels][M]
    std::vector<uint8_t> codes;       // [n_local][cstride] decoded codes, by position - base_pos
    std::vector<int64_t> depth_hist;  // nodes per depth (this shard)
};

// Decodes `payload` and builds the program for shard `rank` of `n_ranks`.
// Returns empty string on success, else an error message.
std::string compile_program(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M, int K,
                            int rank, int n_ranks, int chunk_nodes, ScanProgram* out, int engine = 0);
// engine: 0 = v2 when the shape allows it, 1 = always the first-generation op program
inline bool v2_shape_ok(int M, int K) { return M >= 1 && M <= 16 && K >= 1 && K <= 256; }

}  // namespace dpq
