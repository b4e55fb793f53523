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
// Second-generation layout ("v2", used when M <= 8 and the table has <= 2048 rows): ONE
// fixed 16-byte record per node, eight 16-bit table-row fields, executed by a quarter-warp
// ("strand") that serves 56 queries with 128-bit table loads.
//
//   field  = row * 7            row = m*K + centroid; the table row is 7 x 16 bytes
//   x = plus0 | plus1 << 16     y = plus2 | plus3 << 16
//   z = minus0 | minus1 << 16   w = minus2 | minus3 << 16
//   x bit 14 (ABS)   : 1: dist = sum of all eight rows (full M-term lookup, rows of m = 0..7;
//                         unused fields point at the all-zero row M*K)
//                      0: dist = parent + (plus rows) - (minus rows): one (new, old) pair per
//                         changed subspace, at most four, unused pairs are (row 0, row 0)
//   x bit 15 (CHILD) : the strand's parent register takes this node's distance
//
// A node is delta-encoded when its edge changes <= 4 subspaces AND its parent's distance is
// in the parent register (the node is the first child of the previous node, or a later child
// with only leaves in between); otherwise it gets the full M-term record, which costs the
// same eight table reads and needs no depth stack.  Chunks are v2_chunk_nodes consecutive
// positions, each starting with a full record, so chunks are independent.
//
// Two shapes share this design:
//   narrow (M <= 8,  M*K <= 2048): 8 fields = 16-byte record, 7 x 16-byte table rows (56 queries
//           per CTA), a strand is a quarter warp (8 lanes, 7 active): 4 nodes per warp step;
//   wide   (M <= 16, M*K <= 4096): 16 fields = 32-byte record (plus fields first, minus fields
//           second, <= 8 changed subspaces for a delta record), 3 x 16-byte rows (24 queries per
//           CTA), a strand is 4 lanes (3 active): 8 nodes per warp step.
// In both the table fills shared memory: 2048 x 112 B = 229,376 B, 4096 x 48 B = 196,608 B.
constexpr uint32_t V2_ABS = 1u << 14;
constexpr uint32_t V2_CHILD = 1u << 15;
struct V2Shape {
    int nf;         // fields per record (8 or 16); record = 2*nf bytes
    int lpg;        // active 16-byte lanes per strand = table row bytes / 16
    int sw;         // lanes per strand (8 or 4)
    int rows;       // table rows allocated (2048 or 4096)
    int qb() const { return lpg * 8; }            // queries per CTA
    int row_bytes() const { return lpg * 16; }
    int lut_bytes() const { return rows * row_bytes(); }
    int spw() const { return 32 / sw; }           // strands (= chunks in flight) per warp
    int rec_words() const { return nf / 2; }
};
inline V2Shape v2_shape(int M, int K) {
    if (M <= 8 && M * K <= 2048) return V2Shape{8, 7, 8, 2048};
    return V2Shape{16, 3, 4, 4096};
}

struct ChunkDesc2 {
    uint32_t rec_begin;  // first record of this chunk (records of a chunk are contiguous)
    uint32_t n_nodes;    // records in the chunk (== v2_chunk_nodes except the last)
    uint32_t first_pos;  // global DFS position of the first record
    uint32_t pad;
};

struct ScanProgram {
    int M = 0, K = 0;
    bool v2 = false;
    int v2_chunk_nodes = 64;
    V2Shape shape{8, 7, 8, 2048};
    std::vector<uint32_t> recs;        // v2 records, shape.rec_words() words each
    std::vector<ChunkDesc2> chunks2;
    int64_t v2_delta_nodes = 0;        // nodes that got a delta record
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
    std::vector<uint8_t> codes;       // [n_local][M] decoded codes, by position - base_pos
    std::vector<int64_t> depth_hist;  // nodes per depth (this shard)
};

// Decodes `payload` and builds the program for shard `rank` of `n_ranks`.
// Returns empty string on success, else an error message.
std::string compile_program(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M, int K,
                            int rank, int n_ranks, int chunk_nodes, ScanProgram* out, int engine = 0);
// engine: 0 = v2 when the shape allows it, 1 = always the first-generation op program
// a full record needs either all nf fields used (M == nf) or a spare all-zero table row M*K
inline bool v2_shape_ok(int M, int K) {
    if (M > 16 || M * K > 4096) return false;
    const V2Shape sh = v2_shape(M, K);
    return M == sh.nf || M * K < sh.rows;
}

}  // namespace dpq
