// TEST INFRASTRUCTURE ONLY.  Include-the-header harness around the UNMODIFIED reference.
//
// Compiled by oracle/Makefile (only when /root/reference is present) into
// oracle/_ref/libref_harness.so.  It #includes the reference's own headers where they
// lie, defines the five globals the reference's main() normally owns
// (deltapq_approx_tree_main.cpp:8-12) and exposes, through a plain C ABI, the reference
// functions that have no file output of their own:
//   * query_processing_scan_compressed_codes_opt_in_memory   (deltapq_create_approx_tree.h:3731)
//   * the ADC table it leaves in the global m_sub_distances    (deltapq_create_approx_tree.h:3750-3758)
//   * PQTree::EncodePlain                                      (pq_tree.cpp:192-253)
// (The ground-truth brute force of main.cpp:138-166 lives in the TU that owns main(); it is pinned
// through the reference BINARY instead: tests/golden/make_golden_gt.py runs `pqtree -task groundtruth`
// and commits what it wrote.)
// Nothing here is product code; only tests/, smoke() and bench.py's CPU-baseline leg use it.
#include "pq_tree.h"
#include "utils.h"
#include "deltapq_create_approx_tree.h"

#include <chrono>
#include <sstream>

int PQ_M;
int PQ_K;
int with_id = 0;
string ext = "fvecs";
int dim = 128;

namespace {

std::vector<PQ::Array> to_codewords(const float* cw, int M, int K, int Ds) {
    std::vector<PQ::Array> out(M, PQ::Array(K, std::vector<float>(Ds)));
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k)
            for (int d = 0; d < Ds; ++d) out[m][k][d] = cw[((size_t)m * K + k) * Ds + d];
    return out;
}

uchar** make_decoder() {  // same table as deltapq_approx_tree_main.cpp:312-325
    uchar** decoder = new uchar*[256];
    for (int i = 0; i < 256; i++) {
        std::vector<uchar> slots;
        for (uchar j = 0; j < 8; j++)
            if ((i >> j) & 1) slots.push_back(j);
        decoder[i] = new uchar[slots.size() + 1];
        decoder[i][0] = (uchar)slots.size();
        for (size_t j = 0; j < slots.size(); j++) decoder[i][j + 1] = slots[j];
    }
    return decoder;
}

struct CoutSilencer {  // the reference prints per query; keep test logs readable
    std::streambuf* old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

// Runs the reference in-memory DeltaTree scan for Q queries.  `payload` is the stream
// WITHOUT its 16-byte header (as dmain:630-632 reads it).  Outputs are [Q][topk]
// (DFS position, float distance) ascending, exactly what the reference stores in
// `results`.  If lut_out != NULL it receives the reference's ADC table of every query,
// [Q][M][K].  Returns elapsed seconds for the Q calls (single thread, like dmain:697-705).
double ref_scan_in_memory(const unsigned char* payload, long long n_bytes, long long n_codes,
                          int M, int K, int Ds, const float* cw, const float* queries, int Q,
                          int topk, int* out_pos, float* out_dist, float* lut_out) {
    PQ_M = M;
    PQ_K = K;
    CoutSilencer quiet;
    std::vector<PQ::Array> codewords = to_codewords(cw, M, K, Ds);
    static uchar** decoder = make_decoder();
    std::vector<float> query(M * Ds);
    std::vector<std::pair<int, float> > results(topk);
    auto t0 = std::chrono::steady_clock::now();
    for (int q = 0; q < Q; ++q) {
        query.assign(queries + (size_t)q * M * Ds, queries + (size_t)(q + 1) * M * Ds);
        query_processing_scan_compressed_codes_opt_in_memory(
            const_cast<unsigned char*>(payload), n_bytes, query, topk, M, K, Ds, (uint)n_codes,
            codewords, results, decoder);
        for (int i = 0; i < topk; ++i) {
            out_pos[(size_t)q * topk + i] = results[i].first;
            out_dist[(size_t)q * topk + i] = results[i].second;
        }
        if (lut_out)
            for (int m = 0; m < M; ++m)
                memcpy(lut_out + ((size_t)q * M + m) * K, m_sub_distances[m], sizeof(float) * K);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// PQTree::EncodePlain over n vectors of dimension D (<= M*Ds; zero padded by the reference).
double ref_encode(const float* cw, int M, int K, int Ds, const float* x, long long n, int D,
                  unsigned char* codes) {
    PQ_M = M;
    PQ_K = K;
    CoutSilencer quiet;
    PQTree pqtree(to_codewords(cw, M, K, Ds));
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for
    for (long long i = 0; i < n; ++i) {
        std::vector<float> v(x + (size_t)i * D, x + (size_t)(i + 1) * D);
        std::vector<uchar> c = pqtree.EncodePlain(v);
        for (int m = 0; m < M; ++m) codes[(size_t)i * M + m] = c[m];
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
