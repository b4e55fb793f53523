// Shared host-side pieces of the drop-in command line tools (pqtree, deltapq): the reference's
// flag parsing and on-disk formats.  All compute goes through the C ABI
// (include/dpq.h); nothing here has a CPU compute path.
#pragma once
#include <sys/time.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/dpq.h"

namespace cli {

inline double now_s() {
    timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

// Linear argv scan, value = argv[i+1]; unknown flags are ignored (dmain:26-70, pmain:183-233).
struct Args {
    int argc;
    char** argv;
    std::string str(const char* flag, const std::string& dflt) const {
        std::string v = dflt;
        for (int i = 0; i + 1 < argc; ++i)
            if (std::string(argv[i]) == flag) v = argv[i + 1];
        return v;
    }
    long long num(const char* flag, long long dflt) const {
        long long v = dflt;
        for (int i = 0; i + 1 < argc; ++i)
            if (std::string(argv[i]) == flag) v = atoll(argv[i + 1]);
        return v;
    }
    bool has(const char* flag) const {
        for (int i = 0; i < argc; ++i)
            if (std::string(argv[i]) == flag) return true;
        return false;
    }
};

inline int die(const std::string& msg) {
    std::cerr << msg << std::endl;
    return 1;
}
#define DPQ_TRY(call)                                                        \
    do {                                                                     \
        if ((call) != DPQ_OK) return cli::die(std::string(#call) + ": " + dpq_last_error()); \
    } while (0)

// fvecs / bvecs (utils.cpp:14-71): per vector int32 D then D float32 / D uint8 (widened).
// Bulk reader: up to max_n vectors starting at vector index `first`, row-major floats.
struct VecFile {
    FILE* f = nullptr;
    std::string ext;
    int D = 0;
    long long rec_bytes = 0;
    bool open(const std::string& path, const std::string& ext_) {
        ext = ext_;
        f = fopen(path.c_str(), "rb");
        if (!f) return false;
        int32_t d = 0;
        if (fread(&d, 4, 1, f) != 1) return false;
        D = d;
        rec_bytes = 4 + (long long)D * (ext == "bvecs" ? 1 : 4);
        fseeko(f, 0, SEEK_SET);
        return D > 0;
    }
    // appends to out; returns vectors read
    long long read(long long max_n, std::vector<float>& out) {
        std::vector<unsigned char> raw((size_t)std::min<long long>(max_n, 65536) * rec_bytes);
        long long got_total = 0;
        while (got_total < max_n) {
            long long want = std::min<long long>(max_n - got_total, 65536);
            size_t got = fread(raw.data(), (size_t)rec_bytes, (size_t)want, f);
            if (got == 0) break;
            size_t o = out.size();
            out.resize(o + got * D);
            for (size_t i = 0; i < got; ++i) {
                const unsigned char* r = raw.data() + i * rec_bytes + 4;
                if (ext == "bvecs")
                    for (int d = 0; d < D; ++d) out[o + i * D + d] = (float)r[d];
                else
                    memcpy(&out[o + i * D], r, (size_t)D * 4);
            }
            got_total += (long long)got;
            if ((long long)got < want) break;
        }
        return got_total;
    }
    ~VecFile() {
        if (f) fclose(f);
    }
};

// Codebook text (pq.cpp:267-312): "M,Ks,Ds", per m "m:" then Ks lines of Ds "value," items,
// default ostream float formatting (6 significant digits).
struct Codebook {
    int M = 0, K = 0, Ds = 0;
    std::vector<float> cw;  // [M][K][Ds]
};
inline bool read_codebook(const std::string& path, Codebook& cb) {
    std::ifstream ifs(path);
    if (!ifs.is_open()) return false;
    char c1, c2;
    int v;
    ifs >> cb.M >> c1 >> cb.K >> c2 >> cb.Ds;
    if (!ifs || cb.M < 1 || cb.K < 1 || cb.Ds < 1) return false;
    cb.cw.assign((size_t)cb.M * cb.K * cb.Ds, 0.f);
    for (int m = 0; m < cb.M; ++m) {
        ifs >> v >> c1;
        if (v != m) return false;
        for (int k = 0; k < cb.K; ++k)
            for (int d = 0; d < cb.Ds; ++d) ifs >> cb.cw[((size_t)m * cb.K + k) * cb.Ds + d] >> c1;
    }
    return (bool)ifs;
}
inline bool write_codebook(const std::string& path, const Codebook& cb) {
    std::ofstream ofs(path);
    if (!ofs.is_open()) return false;
    ofs << cb.M << "," << cb.K << "," << cb.Ds << std::endl;
    for (int m = 0; m < cb.M; ++m) {
        ofs << m << ":\n";
        for (int k = 0; k < cb.K; ++k) {
            for (int d = 0; d < cb.Ds; ++d) ofs << cb.cw[((size_t)m * cb.K + k) * cb.Ds + d] << ",";
            ofs << "\n";
        }
    }
    return (bool)ofs;
}

// Codes file (pq_tree.cpp:1011-1081): int64 N then N*M bytes.
inline bool read_codes(const std::string& path, int M, std::vector<uint8_t>& codes, long long& n) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    int64_t nn = 0;
    bool ok = fread(&nn, 8, 1, f) == 1 && nn >= 0;
    if (ok) {
        codes.resize((size_t)nn * M);
        ok = fread(codes.data(), 1, codes.size(), f) == codes.size();
    }
    fclose(f);
    n = nn;
    return ok;
}
inline bool write_file(const std::string& path, const void* hdr, size_t hdr_bytes, const void* body, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = (hdr_bytes == 0 || fwrite(hdr, 1, hdr_bytes, f) == hdr_bytes) &&
              (bytes == 0 || fwrite(body, 1, bytes, f) == bytes);
    fclose(f);
    return ok;
}
inline bool file_exists(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (f) fclose(f);
    return f != nullptr;
}
inline std::string method_suffix(int method) { return method == 2 ? "_WOH" : ""; }

}  // namespace cli
