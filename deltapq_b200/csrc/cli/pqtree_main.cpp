// pqtree -- drop-in for the reference's `pqtree` binary (main.cpp) for the tasks
// learn / encode / groundtruth, same flags (-dataset -ext -task -m -k -N -query_size -topk
// -train_size) and the same files.  Compute runs on the GPU through libdpq.
#include <algorithm>
#include <cfloat>
#include <random>

#include "cli_common.hpp"

using namespace cli;

// Per-subspace k-means (reference: PQ::Learn -> cv::kmeans, pq.cpp:129-157).  The reference's
// learn is RNG driven (unseeded shuffle, OpenCV theRNG) so there is nothing to be bit-exact
// with; this is k-means++ seeding on the host (fixed seed) and Lloyd iterations whose
// assignment step is the GPU nearest-centroid kernel (dpq_encode).
static int learn(const Args& a, const std::string& dataset, const std::string& ext, int M, int K) {
    long long train_size = a.num("-train_size", -1);
    VecFile vf;
    if (!vf.open(dataset + "/learn." + ext, ext)) return die("cannot open " + dataset + "/learn." + ext);
    std::vector<float> x;
    long long n = vf.read(100000, x);  // pmain:246: load_size = 100000
    const int D = vf.D;
    std::cout << "all vectors " << n << std::endl;
    std::mt19937_64 rng(20201113);
    {  // shuffle, then keep train_size (pmain:262-264)
        std::vector<float> tmp(D);
        for (long long i = n - 1; i > 0; --i) {
            long long j = (long long)(rng() % (uint64_t)(i + 1));
            if (i != j) {
                memcpy(tmp.data(), &x[i * D], D * 4);
                memcpy(&x[i * D], &x[j * D], D * 4);
                memcpy(&x[j * D], tmp.data(), D * 4);
            }
        }
        if (train_size > 0 && train_size < n) n = train_size;
    }
    const int Ds = (D + M - 1) / M;  // zero padded to a multiple of M (pq.cpp:114-124)
    Codebook cb;
    cb.M = M;
    cb.K = K;
    cb.Ds = Ds;
    cb.cw.assign((size_t)M * K * Ds, 0.f);
    auto sub = [&](long long i, int m, int d) -> float {
        int col = m * Ds + d;
        return col < D ? x[i * D + col] : 0.f;
    };
    double t0 = now_s();
    for (int m = 0; m < M; ++m) {  // k-means++ seeding
        std::vector<double> best((size_t)n, DBL_MAX);
        long long c = (long long)(rng() % (uint64_t)n);
        for (int k = 0; k < K; ++k) {
            for (int d = 0; d < Ds; ++d) cb.cw[((size_t)m * K + k) * Ds + d] = sub(c, m, d);
            double total = 0;
            for (long long i = 0; i < n; ++i) {
                double s = 0;
                for (int d = 0; d < Ds; ++d) {
                    double df = sub(i, m, d) - cb.cw[((size_t)m * K + k) * Ds + d];
                    s += df * df;
                }
                if (s < best[i]) best[i] = s;
                total += best[i];
            }
            double r = std::uniform_real_distribution<double>(0, total)(rng), acc = 0;
            c = n - 1;
            for (long long i = 0; i < n; ++i) {
                acc += best[i];
                if (acc >= r) {
                    c = i;
                    break;
                }
            }
        }
    }
    std::vector<uint8_t> codes((size_t)n * M), prev;
    for (int it = 0; it < 40; ++it) {  // Lloyd; assignment on the GPU
        DPQ_TRY(dpq_encode(cb.cw.data(), M, K, Ds, x.data(), n, D, codes.data()));
        if (codes == prev) break;
        prev = codes;
        std::vector<double> sum((size_t)M * K * Ds, 0.0);
        std::vector<long long> cnt((size_t)M * K, 0);
        for (long long i = 0; i < n; ++i)
            for (int m = 0; m < M; ++m) {
                int k = codes[i * M + m];
                cnt[(size_t)m * K + k]++;
                for (int d = 0; d < Ds; ++d) sum[((size_t)m * K + k) * Ds + d] += sub(i, m, d);
            }
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k)
                if (cnt[(size_t)m * K + k])
                    for (int d = 0; d < Ds; ++d)
                        cb.cw[((size_t)m * K + k) * Ds + d] =
                            (float)(sum[((size_t)m * K + k) * Ds + d] / (double)cnt[(size_t)m * K + k]);
    }
    std::string path = dataset + "/M" + std::to_string(M) + "K" + std::to_string(K) + "codewords.txt";
    if (!write_codebook(path, cb)) return die("cannot write " + path);
    std::cout << "   Learning codebook uses: " << now_s() - t0 << " sec" << std::endl;
    return 0;
}

static int encode(const Args& a, const std::string& dataset, const std::string& ext, int M, int K) {
    Codebook cb;
    std::string cpath = dataset + "/M" + std::to_string(M) + "K" + std::to_string(K) + "codewords.txt";
    if (!read_codebook(cpath, cb) || cb.M != M || cb.K != K) return die("cannot read codebook " + cpath);
    long long N = a.num("-N", -1);
    VecFile vf;
    if (!vf.open(dataset + "/base." + ext, ext)) return die("cannot open " + dataset + "/base." + ext);
    double t0 = now_s();
    std::vector<uint8_t> codes;
    std::vector<float> buf;
    long long done = 0;
    const long long chunk = 1 << 20;
    std::vector<uint8_t> raw;
    const bool bvecs = ext == "bvecs";  // raw records go to the device as they are (dpq_encode_u8)
    while (N < 0 || done < N) {
        buf.clear();
        long long want = N < 0 ? chunk : std::min(chunk, N - done);
        long long got;
        if (bvecs) {
            raw.resize((size_t)want * vf.rec_bytes);
            got = (long long)fread(raw.data(), (size_t)vf.rec_bytes, (size_t)want, vf.f);
        } else {
            got = vf.read(want, buf);
        }
        if (got == 0) break;
        codes.resize((size_t)(done + got) * M);
        if (bvecs)
            DPQ_TRY(dpq_encode_u8(cb.cw.data(), M, K, cb.Ds, raw.data(), got, vf.D, vf.rec_bytes, 4,
                                  codes.data() + (size_t)done * M));
        else
            DPQ_TRY(dpq_encode(cb.cw.data(), M, K, cb.Ds, buf.data(), got, vf.D, codes.data() + (size_t)done * M));
        done += got;
        printf("\r%lld encoded", done);
        fflush(stdout);
    }
    int64_t n64 = done;
    std::string out = dataset + "/codes.bin.plain.M" + std::to_string(M) + "K" + std::to_string(K) + "N" +
                      std::to_string(done);  // pmain:409-410
    if (!write_file(out, &n64, 8, codes.data(), codes.size())) return die("cannot write " + out);
    std::cout << "\nencoding time " << now_s() - t0 << " sec -> " << out << std::endl;
    return 0;
}

static int groundtruth(const Args& a, const std::string& dataset, const std::string& ext) {
    long long N = a.num("-N", -1);
    int query_size = (int)a.num("-query_size", -1), top_k = (int)a.num("-topk", 1);
    if (query_size == -1) {
        std::cout << "Please specify number of queries to run: -query_size " << std::endl;  // pmain:582
        return 0;
    }
    VecFile qf, bf;
    if (!qf.open(dataset + "/query." + ext, ext)) return die("cannot open query file");
    std::vector<float> q;
    long long nq = qf.read(query_size, q);
    if (nq < query_size) return die("query file holds fewer vectors than -query_size");
    if (!bf.open(dataset + "/base." + ext, ext)) return die("cannot open base file");
    double t0 = now_s();
    dpq_gt* st = nullptr;
    DPQ_TRY(dpq_groundtruth_begin(q.data(), query_size, qf.D, top_k, &st));
    long long done = 0;
    std::vector<float> buf;
    while (N < 0 || done < N) {
        buf.clear();
        long long want = N < 0 ? 200000 : std::min<long long>(200000, N - done);
        long long got = bf.read(want, buf);
        if (got == 0) break;
        DPQ_TRY(dpq_groundtruth_chunk(st, buf.data(), got, done));
        done += got;
    }
    std::vector<uint32_t> ids((size_t)query_size * top_k);
    std::vector<float> dist(ids.size());
    DPQ_TRY(dpq_groundtruth_finish(st, ids.data(), dist.data()));
    std::cout << (now_s() - t0) / query_size * 1000 << " [msec/query] " << std::endl;  // pmain:651
    std::string out = dataset + "/groundtruth/N" + std::to_string(done) + "Top" + std::to_string(top_k) + ".txt";
    std::ofstream ofs(out);  // pqbase.cpp:294-312 (the directory must exist, README.md:94-97)
    if (!ofs.is_open()) return die("cannot write " + out + " (does " + dataset + "/groundtruth exist?)");
    ofs << query_size << "," << top_k << "\n";
    for (int i = 0; i < query_size; ++i) {
        for (int t = 0; t < top_k; ++t) ofs << ids[(size_t)i * top_k + t] << "," << dist[(size_t)i * top_k + t] << ",";
        ofs << "\n";
    }
    std::cout << "Groundtruth written to " << out << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    Args a{argc, argv};
    std::string dataset = a.str("-dataset", ""), ext = a.str("-ext", "fvecs"), task = a.str("-task", "learn");
    int M = (int)a.num("-m", 8), K = (int)a.num("-k", 256);  // pmain:13-14 defaults
    if (dataset.empty()) return die("usage: pqtree -dataset DIR -task learn|encode|groundtruth [-ext fvecs|bvecs] [-m M] [-k K] [-N N] [-query_size Q] [-topk k] [-train_size n]");
    if (task == "learn") return learn(a, dataset, ext, M, K);
    if (task == "encode") return encode(a, dataset, ext, M, K);
    if (task == "groundtruth") return groundtruth(a, dataset, ext);
    return die("pqtree: task '" + task + "' is outside the B200 hot-path scope (learn, encode, groundtruth)");
}
