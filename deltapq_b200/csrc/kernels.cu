// Hand-written sm_100a kernels of the DeltaPQ query hot path.
//
//   lut_kernel      ADC tables (reference DCAT.h:3750-3758), exact + quantised/swizzled
//   scan_kernel     DeltaTree delta scan (reference DCAT.h:3786-3882) over the compiled
//                   op program, lanes = queries, fused per-warp top-k' candidate lists
//   select_kernel   merge of the per-warp lists + exact re-score (reference heap,
//                   DCAT.h:3853-3889) + rigorous margin check
//   fallback_*      exact scan for queries whose margin check failed
//   merge_kernel    k-way merge of per-GPU top-k lists after the NCCL gather
//
// Numerics.  The reference accumulates float table entries in double along the tree path;
// its node distance is therefore (to ~1e-16) float(sum of the node's M table entries).
// scan_kernel instead runs the same parent -> child delta recurrence in EXACT integer
// arithmetic on a fixed-point copy of the table (no error accumulates along a path, the
// only error is the <= 0.5 unit rounding of each of the M entries), keeps topk + slack
// candidates per query, and select_kernel re-scores those with the float table in double.
// A query whose exact k-th distance is not separated from the first rejected fixed-point
// distance by the rounding bound is re-run by the exact fallback, so results never depend
// on the fixed-point precision.
#include "kernels.cuh"

#include <cfloat>

namespace dpq {

// ------------------------------------------------------------------------ PTX helpers --
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}

// ------------------------------------------------------------------------ ADC tables ---
// One CTA per (virtual) query; thread = centroid.  acc follows the reference exactly:
// float accumulator, each term the double square of the float difference, rounded back to
// float after every add (DCAT.h:3754-3757: `m_sub_distances[i][j] += pow(float, 2)`).
__device__ __forceinline__ float adc_entry(const float* __restrict__ c, const float* q, int Ds) {
    float acc = 0.0f;
    for (int d = 0; d < Ds; ++d) {
        float diff = __fsub_rn(c[d], q[d]);
        double t = __dmul_rn((double)diff, (double)diff);
        acc = (float)__dadd_rn((double)acc, t);
    }
    return acc;
}

__global__ void __launch_bounds__(256) lut_kernel(const float* __restrict__ cw, int M, int K, int Ds,
                                                  const float* __restrict__ queries, int Q,
                                                  float* __restrict__ lutf, double* __restrict__ scale,
                                                  uint32_t* __restrict__ qlut, ScanGeom g) {
    extern __shared__ float s_q[];  // M*Ds
    __shared__ int s_max[16];
    const int vq = blockIdx.x;  // virtual query (padding lanes of the last group included)
    const int k = threadIdx.x;
    const int rows = 1 << g.rb;
    // destination of this query inside the swizzled group table
    const int grp = vq / g.qpg, sub = vq % g.qpg;
    const int lane = sub % g.qgl, half = sub / g.qgl;
    uint32_t* dst = qlut ? qlut + ((size_t)grp * g.qgl + lane) * rows : nullptr;
    if (vq >= Q) {  // padding: all-zero table so the lane computes harmless zeros
        if (dst)
            for (int r = k; r < rows; r += blockDim.x) {
                if (g.pack == 1) dst[r] = 0;
                else reinterpret_cast<uint16_t*>(dst)[2 * r + half] = 0;
            }
        return;
    }
    for (int i = k; i < M * Ds; i += blockDim.x) s_q[i] = queries[(size_t)vq * M * Ds + i];
    if (k < 16) s_max[k] = 0;
    __syncthreads();
    float* out = lutf + (size_t)vq * M * K;
    for (int m = 0; m < M; ++m) {
        float v = 0.0f;
        if (k < K) {
            v = adc_entry(cw + ((size_t)m * K + k) * Ds, s_q + m * Ds, Ds);
            out[m * K + k] = v;
        }
        // v >= 0 so integer order == float order
        int vi = __float_as_int(v);
        for (int o = 16; o; o >>= 1) vi = max(vi, __shfl_xor_sync(0xffffffffu, vi, o));
        if ((k & 31) == 0) atomicMax(&s_max[m], vi);
    }
    if (!qlut) return;
    __syncthreads();
    double sum = 0.0;
    for (int m = 0; m < M; ++m) sum += (double)__int_as_float(s_max[m]);
    const double qmax = g.pack == 1 ? (double)(1u << 30) : (double)(32767 - 16);
    const double s = sum > 0.0 ? qmax / sum : 1.0;
    if (k == 0) scale[vq] = s;
    // rows beyond M*K are never addressed; entry (m,k) lives at word (row ^ lane)
    if (k < K)
        for (int m = 0; m < M; ++m) {
            uint32_t v = (uint32_t)__double2ll_rn((double)out[m * K + k] * s);
            int r = (m * K + k) ^ lane;
            if (g.pack == 1) dst[r] = v;
            else reinterpret_cast<uint16_t*>(dst)[2 * r + half] = (uint16_t)v;
        }
}

void launch_lut(const float* d_cw, int M, int K, int Ds, const float* d_queries, int Q,
                float* d_lutf, double* d_scale, uint32_t* d_qlut, const ScanGeom& g,
                cudaStream_t st) {
    int nvq = g.n_groups * g.qpg;
    lut_kernel<<<nvq, 256, (size_t)M * Ds * sizeof(float), st>>>(d_cw, M, K, Ds, d_queries, Q, d_lutf,
                                                              d_scale, d_qlut, g);
}
void launch_lut_plain(const float* d_cw, int M, int K, int Ds, const float* d_queries, int Q,
                      float* d_lutf, cudaStream_t st) {
    ScanGeom g{};
    g.rb = 11;
    g.pack = 1;
    g.qgl = 1;
    g.qpg = 1;
    lut_kernel<<<Q, 256, (size_t)M * Ds * sizeof(float), st>>>(d_cw, M, K, Ds, d_queries, Q, d_lutf,
                                                            nullptr, nullptr, g);
}

// ------------------------------------------------------------------------ scan ---------
// CTA = (query group, tree slice).  Shared memory: the group's fixed-point ADC table,
// [lane][row ^ lane] so that the 32 lanes of a warp, all reading the SAME row, hit 32
// different banks; per-warp depth stacks; per-query shared thresholds.
// Each warp pulls chunks of the op program and runs the delta recurrence with lanes =
// queries: one table row read per changed subspace for the old and the new centroid.
template <int PACK>
struct Lists {
    uint64_t* base;  // this warp's lists: [kp][32*PACK]
    int kp;
};

// Sorted insert into one lane's candidate list (global scratch).  Returns new count.
__device__ __forceinline__ int list_insert(uint64_t* L, int stride, int kp, int n, uint64_t key) {
    if (n == kp) {
        if (key >= L[(size_t)(kp - 1) * stride]) return n;
    }
    int i = n < kp ? n : kp - 1;
    while (i > 0) {
        uint64_t prev = L[(size_t)(i - 1) * stride];
        if (prev <= key) break;
        L[(size_t)i * stride] = prev;
        --i;
    }
    L[(size_t)i * stride] = key;
    return n < kp ? n + 1 : n;
}

template <int RB, int PACK>
__global__ void __launch_bounds__(512, 1) scan_kernel(const ScanArgs a) {
    constexpr int ROWS = 1 << RB;
    constexpr int LEVELS = RB == 11 ? 8 : 16;
    constexpr uint32_t FMASK = (uint32_t)(ROWS - 1) << 2;
    constexpr int TSH = RB + 2;
    constexpr int LW = 32 * PACK;
    constexpr uint32_t XINF = PACK == 1 ? kInf31 : 0x8000u;  // exclusive bound "accept all"
    extern __shared__ __align__(128) unsigned char smem[];
    const ScanGeom& g = a.g;
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(smem);
    uint32_t* s_stack = s_lut + (size_t)g.qgl * ROWS;       // [n_warps][LEVELS][32]
    uint32_t* s_thr = s_stack + (size_t)g.n_warps * LEVELS * 32;  // [LW]
    int* s_next = reinterpret_cast<int*>(s_thr + LW);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_next + 2);

    const int item = blockIdx.x;
    const int slice = item / g.n_groups, grp = item % g.n_groups;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this slice's chunk range
    const int c_lo = (int)((int64_t)a.n_chunks * slice / g.n_slices);
    const int c_hi = (int)((int64_t)a.n_chunks * (slice + 1) / g.n_slices);

    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        *s_next = c_lo;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < LW; i += blockDim.x) s_thr[i] = XINF;
    __syncthreads();
    if (threadIdx.x == 0) {  // TMA bulk copy of the group's table: qgl regions of ROWS*4 bytes
        const uint32_t bytes = (uint32_t)g.qgl * ROWS * 4u;
        mbar_expect_tx(s_bar, bytes);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.qlut) + (size_t)grp * bytes;
        for (uint32_t o = 0; o < bytes; o += ROWS * 4u) bulk_g2s(smem + o, src + o, ROWS * 4u, s_bar);
    }
    mbar_wait(s_bar, 0);

    const int ql = lane < g.qgl ? lane : g.qgl - 1;  // idle lanes alias the last region
    // a (lane, half) is valid when it maps to a real query of this group
    bool valid[PACK];
    uint32_t livemask = 0;
#pragma unroll
    for (int h = 0; h < PACK; ++h) {
        valid[h] = lane < g.qgl && grp * g.qpg + h * g.qgl + lane < a.Q;
        if (valid[h]) livemask |= PACK == 1 ? 1u : (0x8000u << (16 * h));
    }
    const bool live = livemask != 0;
    const uint32_t Lx = ((uint32_t)ql << (RB + 2)) | ((uint32_t)ql << 2);
    const unsigned char* lutb = smem;
    uint32_t* stack = s_stack + (size_t)warp * LEVELS * 32 + lane;
#define DPQ_LUT(off) (*reinterpret_cast<const uint32_t*>(lutb + (off)))

    // candidate lists of this warp
    uint64_t* lists = a.cand + ((size_t)item * g.n_warps + warp) * (size_t)g.kp * LW;
    const int kp = g.kp;
    int cnt[PACK];
    uint32_t own[PACK];  // own exclusive bound (kp-th distance + 1 once the list is full)
#pragma unroll
    for (int h = 0; h < PACK; ++h) {
        cnt[h] = 0;
        own[h] = XINF;
    }
    uint32_t thr = 0;  // PACK 1: exclusive bound; PACK 2: packed ((X-1)|0x8000) per half

    auto refresh = [&]() {
        if (PACK == 1) {
            thr = live ? min(own[0], s_thr[lane]) : 0u;  // bound 0 rejects everything
        } else {
            uint32_t x0 = min(own[0], s_thr[lane]);
            uint32_t x1 = min(own[PACK - 1], s_thr[lane + 32 * (PACK - 1)]);
            thr = ((x0 - 1u) | 0x8000u) | (((x1 - 1u) | 0x8000u) << 16);
        }
    };
    // slow path: at least one lane has a candidate for node `pos` with packed distance d
    auto offer = [&](uint32_t d, uint32_t pos) {
        if (live) {
#pragma unroll
            for (int h = 0; h < PACK; ++h) {
                uint32_t dh = PACK == 1 ? d : ((d >> (16 * h)) & 0xFFFFu);
                uint32_t x = min(own[h], s_thr[lane + 32 * h]);
                if (dh < x && valid[h]) {
                    uint64_t key = ((uint64_t)dh << 32) | pos;
                    cnt[h] = list_insert(lists + lane + 32 * h, LW, kp, cnt[h], key);
                    if (cnt[h] == kp) {
                        uint32_t kth = (uint32_t)(lists[(size_t)(kp - 1) * LW + lane + 32 * h] >> 32);
                        own[h] = kth + 1u;
                        atomicMin(&s_thr[lane + 32 * h], kth + 1u);
                    }
                }
            }
        }
        refresh();
    };
    auto is_cand = [&](uint32_t d) -> bool {
        if (PACK == 1) return d < thr;
        return ((thr - d) & livemask) != 0u;
    };

    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(s_next, 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= c_hi) break;
        const ChunkDesc cd = a.chunks[c];
        const int n_anc = (int)(cd.n_anc_flags & 0xFFu);
        refresh();
        // chunk prologue: full M-term sums for the ancestors of the first node
        uint32_t par = 0;
        {
            const uint8_t* anc = a.anc + (size_t)c * LEVELS * g.M;
            for (int lev = 0; lev < n_anc; ++lev) {
                uint32_t d = 0;
                for (int m = 0; m < g.M; ++m) {
                    uint32_t row4 = (uint32_t)(m * g.K + anc[lev * g.M + m]) << 2;
                    d += DPQ_LUT(row4 ^ Lx);
                }
                stack[lev * 32] = d;
                par = d;
                if (lev == 0 && (cd.n_anc_flags & CHUNK_EMIT_ROOT)) {
                    if (__any_sync(0xffffffffu, is_cand(d))) offer(d, 0u);
                }
            }
        }
        uint32_t acc = par;
        uint32_t pos = cd.first_pos;
        const uint4* qp = a.ops + cd.quad_begin;
        const uint4* qe = qp + cd.n_quads;
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (qp < qe) nxt = __ldg(qp);
        while (qp < qe) {
            const uint4 cur = nxt;
            ++qp;
            if (qp < qe) nxt = __ldg(qp);
#define DPQ_OP(w)                                                                     \
    {                                                                                 \
        const uint32_t fo = ((w) & FMASK) ^ Lx;                                       \
        const uint32_t to = (((w) >> TSH) & FMASK) ^ Lx;                              \
        acc = acc + DPQ_LUT(to) - DPQ_LUT(fo);                                        \
        if ((int)(w) < 0) {                                                           \
            const uint32_t d = acc;                                                   \
            if (__any_sync(0xffffffffu, is_cand(d))) offer(d, pos);                   \
            if ((w) & (OP_CHILD | OP_AUX)) {                                          \
                const uint32_t lev = RB == 11 ? (((w) >> 26) & 7u)                    \
                                              : (((w) & 3u) | ((((w) >> 14) & 3u) << 2)); \
                if ((w) & OP_CHILD) {                                                 \
                    par = d;                                                          \
                    if ((w) & OP_AUX) stack[lev * 32] = d;                            \
                } else {                                                              \
                    par = stack[lev * 32];                                            \
                }                                                                     \
            }                                                                         \
            acc = par;                                                                \
            ++pos;                                                                    \
        }                                                                             \
    }
            DPQ_OP(cur.x)
            DPQ_OP(cur.y)
            DPQ_OP(cur.z)
            DPQ_OP(cur.w)
#undef DPQ_OP
        }
    }
#undef DPQ_LUT
#pragma unroll
    for (int h = 0; h < PACK; ++h)
        a.cand_cnt[((size_t)item * g.n_warps + warp) * LW + lane + 32 * h] = live ? (uint32_t)cnt[h] : 0u;
}

cudaError_t launch_scan(const ScanArgs& a, cudaStream_t st) {
    const ScanGeom& g = a.g;
    dim3 grid((unsigned)(g.n_groups * g.n_slices)), block((unsigned)(g.n_warps * 32));
    void (*k)(const ScanArgs) = nullptr;
    if (g.rb == 11 && g.pack == 1) k = scan_kernel<11, 1>;
    else if (g.rb == 11 && g.pack == 2) k = scan_kernel<11, 2>;
    else if (g.rb == 12 && g.pack == 1) k = scan_kernel<12, 1>;
    else if (g.rb == 12 && g.pack == 2) k = scan_kernel<12, 2>;
    else return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, block, g.smem_bytes, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------ select -------
// One warp per query: k'-way merge of the per-(slice, warp) sorted candidate lists, exact
// re-score of the k' winners, sort by (exact distance, position), margin check.
constexpr int SEL_WARPS = 4;
constexpr int SEL_MAXKP = 256;
constexpr int SEL_MAXLISTS_PER_LANE = 64;

__global__ void __launch_bounds__(SEL_WARPS * 32) select_kernel(const SelectArgs a) {
    __shared__ uint64_t s_sel[SEL_WARPS][SEL_MAXKP];
    __shared__ uint64_t s_exact[SEL_WARPS][SEL_MAXKP];
    const ScanGeom& g = a.g;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * SEL_WARPS + w;
    if (q >= a.Q) return;
    const int LW = 32 * g.pack;
    const int grp = q / g.qpg, sub = q % g.qpg;
    const int sl = (sub % g.qgl) + 32 * (sub / g.qgl);
    const int n_lists = g.n_slices * g.n_warps;
    const int kp = g.kp;

    // cursors over this lane's lists (list id = lane + 32*j)
    uint16_t cur[SEL_MAXLISTS_PER_LANE];
    uint16_t len[SEL_MAXLISTS_PER_LANE];
    const int my = (n_lists - lane + 31) / 32;
    auto list_base = [&](int id) -> const uint64_t* {
        int s = id / g.n_warps, ww = id % g.n_warps;
        size_t item = (size_t)s * g.n_groups + grp;
        return a.cand + ((item * g.n_warps + ww) * (size_t)kp) * LW + sl;
    };
    for (int j = 0; j < my; ++j) {
        int id = lane + 32 * j;
        int s = id / g.n_warps, ww = id % g.n_warps;
        size_t item = (size_t)s * g.n_groups + grp;
        len[j] = (uint16_t)a.cand_cnt[(item * g.n_warps + ww) * LW + sl];
        cur[j] = 0;
    }
    int n_sel = 0;
    for (int r = 0; r < kp; ++r) {
        uint64_t best = ~0ull;
        int bj = -1;
        for (int j = 0; j < my; ++j)
            if (cur[j] < len[j]) {
                uint64_t k = list_base(lane + 32 * j)[(size_t)cur[j] * LW];
                if (k < best) {
                    best = k;
                    bj = j;
                }
            }
        uint64_t m = best;
        for (int o = 16; o; o >>= 1) {
            uint64_t t = __shfl_xor_sync(0xffffffffu, m, o);
            m = t < m ? t : m;
        }
        if (m == ~0ull) break;
        if (best == m && bj >= 0) cur[bj]++;  // keys are unique (position is part of the key)
        if (lane == 0) s_sel[w][r] = m;
        n_sel = r + 1;
    }
    __syncwarp();
    // exact re-score: float(sum in double of the node's M float table entries)
    const float* lut = a.lutf + (size_t)q * g.M * g.K;
    for (int j = lane; j < n_sel; j += 32) {
        uint32_t pos = (uint32_t)s_sel[w][j];
        const uint8_t* code = a.codes + (size_t)((int64_t)pos - a.base_pos) * g.M;
        double d = 0.0;
        for (int m = 0; m < g.M; ++m) d += (double)lut[m * g.K + code[m]];
        s_exact[w][j] = ((uint64_t)__float_as_uint((float)d) << 32) | pos;
    }
    __syncwarp();
    // rank sort (keys unique) and output
    uint64_t kth_exact = 0;
    for (int j = lane; j < n_sel; j += 32) {
        uint64_t k = s_exact[w][j];
        int rank = 0;
        for (int i = 0; i < n_sel; ++i) rank += s_exact[w][i] < k;
        if (rank < a.topk) a.out_key[(size_t)q * a.topk + rank] = k;
        if (rank == a.topk - 1) kth_exact = k;
    }
    for (int j = n_sel + lane; j < a.topk; j += 32)
        a.out_key[(size_t)q * a.topk + j] = ((uint64_t)__float_as_uint(FLT_MAX) << 32) | 0xFFFFFFFFull;
    for (int o = 16; o; o >>= 1) {
        uint64_t t = __shfl_xor_sync(0xffffffffu, kth_exact, o);
        kth_exact = t > kth_exact ? t : kth_exact;
    }
    if (lane == 0) {
        // Every node NOT among the k' winners has fixed-point distance >= G (the k'-th).
        // Its exact distance is >= (G - E) / scale with E = M/2 + 2 rounding units.
        uint32_t flag = 0;
        float bound = FLT_MAX;
        if (n_sel >= a.topk) bound = __uint_as_float((uint32_t)(kth_exact >> 32));
        if (n_sel == kp) {
            double G = (double)(uint32_t)(s_sel[w][kp - 1] >> 32);
            double E = 0.5 * g.M + 2.0;
            if (!((double)bound * a.scale[q] + E < G)) flag = 1;
        }
        if (a.force_fallback) flag = 1;
        a.bound[q] = bound;
        if (flag) {
            uint32_t slot = atomicAdd(a.n_flagged, 1u);
            if (slot < (uint32_t)a.max_flagged) a.flagged[slot] = (uint32_t)q;
        }
    }
}

void launch_select(const SelectArgs& a, cudaStream_t st) {
    int blocks = (a.Q + SEL_WARPS - 1) / SEL_WARPS;
    select_kernel<<<blocks, SEL_WARPS * 32, 0, st>>>(a);
}

// ------------------------------------------------------------------------ fallback -----
// Exact plain ADC scan over the decoded codes for the (rare) flagged queries.
__global__ void __launch_bounds__(256) fallback_collect_kernel(const FallbackArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int n_flagged = min((int)*a.n_flagged, a.max_flagged);
    for (int f = 0; f < n_flagged; ++f) {
        const uint32_t q = a.flagged[f];
        const float* lut = a.lutf + (size_t)q * a.M * a.K;
        const float bound = a.bound[q];
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_local; i += stride) {
            const uint8_t* code = a.codes + (size_t)i * a.M;
            double d = 0.0;
            for (int m = 0; m < a.M; ++m) d += (double)lut[m * a.K + code[m]];
            float df = (float)d;
            if (df <= bound) {
                uint32_t slot = atomicAdd(&a.buf_cnt[f], 1u);
                if (slot < (uint32_t)a.cap)
                    a.buf[(size_t)f * a.cap + slot] =
                        ((uint64_t)__float_as_uint(df) << 32) | (uint32_t)(i + a.base_pos);
                else
                    *a.overflow = 1u;
            }
        }
    }
}

__global__ void __launch_bounds__(256) fallback_finish_kernel(const FallbackArgs a) {
    extern __shared__ uint64_t s_keys[];  // cap
    const int n_flagged = min((int)*a.n_flagged, a.max_flagged);
    if (threadIdx.x == 0 && blockIdx.x == 0 && (int)*a.n_flagged > a.max_flagged) *a.overflow = 2u;
    for (int f = blockIdx.x; f < n_flagged; f += gridDim.x) {
    const uint32_t q = a.flagged[f];
    int n = min((int)a.buf_cnt[f], a.cap);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        s_keys[i] = i < n ? a.buf[(size_t)f * a.cap + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint64_t x = s_keys[i], y = s_keys[ixj];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        s_keys[i] = y;
                        s_keys[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < a.topk; i += blockDim.x)
        a.out_key[(size_t)q * a.topk + i] =
            i < n ? s_keys[i] : (((uint64_t)__float_as_uint(FLT_MAX) << 32) | 0xFFFFFFFFull);
    __syncthreads();
    }
}

void launch_fallback(const FallbackArgs& a, cudaStream_t st) {
    // Always enqueued; both kernels read the flagged count on the device and exit at once
    // when it is zero, so the common path needs no host round trip.
    fallback_collect_kernel<<<148 * 4, 256, 0, st>>>(a);
    int blocks = a.max_flagged < 296 ? a.max_flagged : 296;
    fallback_finish_kernel<<<blocks, 256, (size_t)a.cap * sizeof(uint64_t), st>>>(a);
}

// ------------------------------------------------------------------------ merge --------
__global__ void merge_kernel(const uint64_t* __restrict__ keys, int n_lists, int Q, int topk,
                             uint64_t* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    int cur[64];
    for (int l = 0; l < n_lists; ++l) cur[l] = 0;
    for (int r = 0; r < topk; ++r) {
        uint64_t best = ~0ull;
        int bl = -1;
        for (int l = 0; l < n_lists; ++l)
            if (cur[l] < topk) {
                uint64_t k = keys[((size_t)l * Q + q) * topk + cur[l]];
                if (k < best) {
                    best = k;
                    bl = l;
                }
            }
        if (bl >= 0) cur[bl]++;
        out[(size_t)q * topk + r] =
            bl >= 0 ? best : (((uint64_t)__float_as_uint(FLT_MAX) << 32) | 0xFFFFFFFFull);
    }
}

void launch_merge(const uint64_t* d_keys, int n_lists, int Q, int topk, uint64_t* d_out,
                  cudaStream_t st) {
    merge_kernel<<<(Q + 127) / 128, 128, 0, st>>>(d_keys, n_lists, Q, topk, d_out);
}

}  // namespace dpq
