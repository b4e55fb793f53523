/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the DeltaPQ query hot path.
 *
 * Plain C11 restatement of the reference algorithms (RunhuiWang/DeltaPQ); every function
 * cites the reference file:line it follows.  Short names: DCAT.h =
 * deltapq_create_approx_tree.h, dmain = deltapq_approx_tree_main.cpp, pmain = main.cpp,
 * CT.h = create_tree.h.
 *
 * Pinning: the reference ships no golden vectors.  This oracle is
 * pinned against outputs of the UNMODIFIED reference compiled in oracle/_ref (see
 * oracle/Makefile): tests/golden/ holds reference-generated fixtures + the generating
 * script (tests/golden/make_golden.py), and tests/test_oracle_vs_ref.py re-runs the
 * comparison live whenever oracle/_ref is present.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library.  The product (deltapq_b200/csrc, include/dpq.h) never does.
 */
#ifndef DPQ_ORACLE_H
#define DPQ_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* QNode file record, M = 8 layout (DCAT.h:79-101; sizeof == 60). */
#pragma pack(push, 1)
typedef struct {
    uint32_t vec_id, parent_pos, child_pos_start, child_num, sub_tree_size;
    float qdist, max_dist, max_dist2p;
    uint8_t diff_num, depth;
    uint8_t diffs[8][3]; /* {m, from, to} */
    uint8_t pad[2];
} dpqo_qnode8;
#pragma pack(pop)

/* ADC table of one query: lut[m*K+k]  (DCAT.h:3750-3758 == :2841-2849). */
void dpqo_lut(const float* cw, int M, int K, int Ds, const float* query, float* lut);

/* Single-query in-memory DeltaTree scan (DCAT.h:3731-3890).  payload = stream without the
 * 16-byte header.  Works for M == 8 (reference format) and M > 8 (extension format:
 * ceil(M/8)-byte little-endian bitmap, depth nibbles unmasked).  Results ascending by
 * distance: (DFS position, float distance).  The trailing node of an even-N tree is
 * reported at its true position N-1 (the reference reports N).
 * If node_dist != NULL it receives the float distance of every node, [n_codes].
 * Returns the number of payload bytes consumed. */
int64_t dpqo_scan(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M, int K,
                  const float* lut, int topk, int32_t* out_pos, float* out_dist,
                  float* node_dist);

/* Lossless decode of the stream into per-position codes [n_codes][M], depth [n_codes] and
 * parent position [n_codes] (root: -1).  Any output may be NULL.  Returns bytes consumed. */
int64_t dpqo_decode(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M,
                    uint8_t* codes, uint8_t* depth, int32_t* parent);

/* PQTree::EncodePlain (pq_tree.cpp:192-253): x is [n][D] with D <= M*Ds (zero padded). */
void dpqo_encode(const float* cw, int M, int K, int Ds, const float* x, int64_t n, int D,
                 uint8_t* codes);

/* Canonical (stable-sort) edge search, method 1 or 2 (DCAT.h:1207-1332, 445-627, 629-792).
 * edges: [n_codes-1][2] (parent id, child id) in emission order. Returns #edges. */
int64_t dpqo_find_edges(const uint8_t* codes, int64_t n_codes, int M, int K,
                        int max_height_folds, int method, uint32_t* edges, uint32_t* root_id);

/* K x K centroid distance tables, tables[m][a*K+b] (dmain:101-118). */
void dpqo_centroid_tables(const float* cw, int M, int K, int Ds, float* tables);

/* Edges -> DFS layout (DCAT.h:1334-1487, 1067-1104, 1156-1183).  Outputs per position:
 * vec_id, parent_pos (root 0xFFFFFFFF), child_num, depth, max_dist, max_dist2p.  All
 * arrays [n_codes]; any may be NULL. */
void dpqo_layout(const uint8_t* codes, int64_t n_codes, int M, int K, const uint32_t* edges,
                 uint32_t root_id, const float* tables, uint32_t* vec_id, uint32_t* parent_pos,
                 uint32_t* child_num, uint8_t* depth, float* max_dist, float* max_dist2p);

/* QNode[N+1] array exactly as the reference writes it (M == 8 only). */
void dpqo_qnodes8(const uint8_t* codes, int64_t n_codes, const uint32_t* vec_id,
                  const uint32_t* parent_pos, const uint32_t* child_num, const uint8_t* depth,
                  const float* max_dist, const float* max_dist2p, dpqo_qnode8* nodes);

/* Stream size and writer (DCAT.h:1765-1842; M > 8 uses the extension format).
 * dpqo_stream_bytes returns n_bytes; dpqo_stream fills payload[n_bytes]. */
int64_t dpqo_stream_bytes(const uint8_t* codes, int64_t n_codes, int M, const uint32_t* vec_id,
                          const uint32_t* parent_pos);
int64_t dpqo_stream(const uint8_t* codes, int64_t n_codes, int M, const uint32_t* vec_id,
                    const uint32_t* parent_pos, const uint8_t* depth, uint8_t* payload);

/* Brute-force ground truth (pmain:138-166, 569-669): base [n][D], ids start at id0.
 * State = per-query max-heaps [Q][topk] (dist float, id), count[Q]; call once per chunk,
 * then dpqo_groundtruth_finish to get ascending results. */
void dpqo_groundtruth_chunk(const float* base, int64_t n, int64_t id0, const float* queries,
                            int Q, int D, int topk, float* heap_dist, uint32_t* heap_id,
                            int32_t* count);
void dpqo_groundtruth_finish(int Q, int topk, float* heap_dist, uint32_t* heap_id,
                             const int32_t* count, uint32_t* out_id, float* out_dist);

#ifdef __cplusplus
}
#endif
#endif
