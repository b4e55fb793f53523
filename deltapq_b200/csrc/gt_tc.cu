// Ground truth, tensor-core form: exact brute force = a dense contraction
// ||x - q||^2 = ||x||^2 + ||q||^2 - 2 x.q, so the Q x n x D inner products run on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM) as a FILTER, and only the few candidates that
// can reach a query's top-k are re-scored in the reference's own arithmetic (pmain:150-156: float
// difference, float product, double sum).  The result is therefore exactly the SIMT path's.
//
//   gt_prep_kernel    per vector: ||v||^2 (double sum), rounded conservatively, and ||v||
//   gt_tc_filter      CTA = 128 queries x a range of 256-vector base tiles.  Operands are split
//                     into bf16 hi + lo parts on the fly (x = hi + lo + O(2^-18 x)); per 64-dim
//                     chunk three MMA groups accumulate hi.hi + hi.lo + lo.hi into one 128 x 256
//                     fp32 TMEM tile.  Epilogue: every thread owns one query row, reads its row
//                     with tcgen05.ld and appends (query, id) when
//                         ||x||^2 - 2 dot - err(x, q)  <=  cap_q - ||q||^2,
//                     err = the rigorous bound on everything the split, the fp32 accumulation and
//                     the norms can be off by, cap_q = the query's current exact k-th distance.
//   gt_rescore_kernel one warp per query: exact distance of every candidate, k best by
//                     (distance, id) merged into the running state.
//
// Shared-memory operand layout = the canonical K-major UMMA layout with the 128-byte swizzle: a
// chunk is 64 dims, so a tile row is exactly one 128-byte swizzle row (see smem_desc below).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "gt_tc.cuh"
#include "kernels.cuh"  // smem_u32, mbar_init

namespace dpq {
namespace {

constexpr int TQ = 128;   // queries per CTA = MMA M = TMEM lanes
constexpr int TB = 256;   // base vectors per tile = MMA N = TMEM columns
constexpr int KC = 64;    // dims per chunk (4 MMA k-steps of 16)
constexpr int A_PART = TQ * KC * 2;  // bytes of one bf16 part of the query chunk (16 KB)
constexpr int B_PART = TB * KC * 2;  // 32 KB
constexpr int SMEM_OPER = 2 * A_PART + 2 * B_PART;  // hi + lo of both: 96 KB
constexpr int SMEM_TOTAL = SMEM_OPER + TB * 8 + 64;

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N = 256, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TB >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);

// K-major operand tile with the 128-byte swizzle: row r of the tile is 128 contiguous bytes (the
// chunk's 64 bf16) at r * 128, its 16-byte piece c stored at position c ^ (r & 7); 8-row groups are
// 1024 bytes apart (stride byte offset), the leading byte offset is unused (1).  One MMA k-step
// (16 elements) advances the start address by 32 bytes inside the swizzle atom; tiles are 1024-byte
// aligned (descriptor fields as in cute/arch/mma_sm100_desc.hpp, layout type 2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t phase) {
    for (uint32_t spin = 0; spin < (1u << 27); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
        if (done) return true;
    }
    return false;
}

// 8 consecutive floats -> bf16 hi and lo parts (16 bytes each)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// rows x KC floats of `src` (row stride D, first row row0, n_rows valid, dims d0.. valid below D)
// -> hi / lo operand parts in the canonical layout.  A warp covers 8 rows x 4 k-columns per step:
// 128-byte global segments per row, 512-byte conflict-free shared stores.
constexpr int NT = 256;  // threads per CTA: 8 warps fill; warps w and w + 4 share TMEM lanes 32 (w & 3) ..

// One operand tile of a chunk: ROWS x KC floats of `src` (row stride D, first row row0, n_rows valid,
// dims from d0, valid below D) -> bf16 hi / lo parts in the canonical layout.  A work item is one
// row x 32 dims handled by four lanes (kq = 0..3): lane kq loads the 16-byte pieces kq and kq + 4 of
// the 128-byte segment, so one warp request reads 8 rows x 64 contiguous bytes (every sector fully
// used), and writes each converted piece (4 bf16 = 8 bytes) to its half of a core-matrix row: a
// warp store covers 2 x 128 contiguous bytes, the minimum two wavefronts.  load() issues every
// global load of the thread's items, store() converts and writes them: the caller loads BOTH
// operands before storing either, so one round of memory latency covers the whole chunk.
template <int ROWS>
struct OperandFill {
    static constexpr int ITEMS = ROWS * (KC / 8);
    static constexpr int PER = ITEMS / NT;  // items per thread
    static_assert(ITEMS % NT == 0, "tile shape");
    float4 pa[PER], pb[PER];  // pieces kq and kq + 4

    __device__ __forceinline__ static void coords(int i, int& row, int& kh, int& kq) {
        const int r = i & 7, rest = i >> 5;
        kq = (i >> 3) & 3;
        const int rg = rest % (ROWS / 8);
        kh = rest / (ROWS / 8);  // which 32-dim half of the chunk
        row = rg * 8 + r;
    }
    __device__ __forceinline__ void load(const float* __restrict__ src, int64_t row0, int64_t n_rows, int D, int d0) {
        const bool whole = row0 + ROWS <= n_rows && d0 + KC <= D && (D & 3) == 0;  // 16-byte aligned full tile
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            int row, kh, kq;
            coords(threadIdx.x + u * NT, row, kh, kq);
            const int d = d0 + kh * 32;
            if (whole) {
                const float4* p = reinterpret_cast<const float4*>(src + (size_t)(row0 + row) * D + d);
                pa[u] = __ldg(p + kq);
                pb[u] = __ldg(p + kq + 4);
            } else {  // edge tiles: rows / dims beyond the data are zero
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int dd = d + (j < 4 ? kq * 4 + j : (kq + 4) * 4 + j - 4);
                    v[j] = (row0 + row < n_rows && dd < D) ? __ldg(src + (size_t)(row0 + row) * D + dd) : 0.0f;
                }
                pa[u] = make_float4(v[0], v[1], v[2], v[3]);
                pb[u] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
    __device__ __forceinline__ void store(unsigned char* s_hi, unsigned char* s_lo) const {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            int row, kh, kq;
            coords(threadIdx.x + u * NT, row, kh, kq);
            const float v[8] = {pa[u].x, pa[u].y, pa[u].z, pa[u].w, pb[u].x, pb[u].y, pb[u].z, pb[u].w};
            uint4 hi, lo;
            split8(v, hi, lo);  // .x.y = piece kq, .z.w = piece kq + 4
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int piece = kq + 4 * h, k8 = kh * 4 + (piece >> 1);
                const int off = row * 128 + ((k8 ^ (row & 7)) << 4) + (piece & 1) * 8;
                *reinterpret_cast<uint2*>(s_hi + off) = h ? make_uint2(hi.z, hi.w) : make_uint2(hi.x, hi.y);
                *reinterpret_cast<uint2*>(s_lo + off) = h ? make_uint2(lo.z, lo.w) : make_uint2(lo.x, lo.y);
            }
        }
    }
};

}  // namespace

// ||v||^2 in double; nlo = a float <= the true value (used on the base side), nhi >= it, len >= ||v||
__global__ void gt_prep_kernel(const float* __restrict__ x, int64_t n, int D, float* __restrict__ nlo,
                               float* __restrict__ nhi, float* __restrict__ len) {
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (v >= n) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
        const double t = (double)x[(size_t)v * D + d];
        s += t * t;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        nlo[v] = __double2float_rd(s * (1.0 - 2e-6));
        nhi[v] = __double2float_ru(s * (1.0 + 2e-6));
        len[v] = __double2float_ru(sqrt(s) * (1.0 + 1e-6));
    }
}

// thr[q] = cap_q - ||q||^2 (rounded up), cap_q = current k-th best distance of the query widened by
// the relative slack that covers the reference arithmetic's own rounding; qerr[q] = c_err * ||q||
__global__ void gt_thr_kernel(const unsigned long long* __restrict__ state, int topk, int Q,
                              const float* __restrict__ q_nlo, const float* __restrict__ q_nhi,
                              const float* __restrict__ q_len, float c_err, float* __restrict__ thr,
                              float* __restrict__ qerr) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const float cap = __uint_as_float((uint32_t)(state[(size_t)q * topk + topk - 1] >> 32));
    // d_ref = d_true (1 +- 2^-22) and the epilogue's two fp32 FMAs round by 2^-24 of their operands;
    // d_true <= 2 (||x||^2 + ||q||^2): the slack on ||x||^2 is taken on the base side (nlo = 1 - 2e-6),
    // the one on ||q||^2 here
    const double capd = (double)cap * (1.0 + 1e-6) + 4e-6 * (double)q_nhi[q];
    thr[q] = cap >= FLT_MAX ? FLT_MAX : __double2float_ru(capd - (double)q_nlo[q]);
    qerr[q] = __fmul_ru(c_err, q_len[q]);
}

__global__ void __launch_bounds__(NT, 2) gt_tc_filter_kernel(const GtTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA_hi = smem;
    unsigned char* sA_lo = smem + A_PART;
    unsigned char* sB_hi = smem + 2 * A_PART;
    unsigned char* sB_lo = smem + 2 * A_PART + B_PART;
    float2* s_x = reinterpret_cast<float2*>(smem + SMEM_OPER);  // per column: (||x||^2 low, ||x|| high)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + SMEM_OPER + TB * 8);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int q0 = blockIdx.y * TQ;
    const int lane_grp = warp & 3, col_half = warp >> 2;  // TMEM lanes 32 lane_grp .., columns 128 col_half ..
    const int q = q0 + lane_grp * 32 + (threadIdx.x & 31);  // this thread's query row = TMEM lane
    const int64_t n_tiles = (a.n + TB - 1) / TB;
    const int64_t t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // one warp allocates the 256 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)TB)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;

    const float thr = q < a.Q ? a.thr[q] : -FLT_MAX;
    const float qerr = q < a.Q ? a.qerr[q] : 0.0f;
    const int n_chunks = (a.D + KC - 1) / KC;
    uint32_t phase = 0;
    bool ok = (smem_u32(smem) & 1023u) == 0;  // the swizzle is a function of the absolute address

    for (int64_t t = t_lo; t < t_hi && ok; ++t) {
        const int64_t x0 = t * TB;
        for (int c = threadIdx.x; c < TB; c += blockDim.x)
            s_x[c] = x0 + c < a.n ? make_float2(a.x_nlo[x0 + c], a.x_len[x0 + c]) : make_float2(__int_as_float(0x7f800000), 0.0f);  // +inf: a padded column never passes
        for (int ch = 0; ch < n_chunks && ok; ++ch) {
            {
                OperandFill<TB> fb;
                OperandFill<TQ> fa;
                fb.load(a.base, x0, a.n, a.D, ch * KC);
                fa.load(a.queries, q0, a.Q, a.D, ch * KC);
                fb.store(sB_hi, sB_lo);
                fa.store(sA_hi, sA_lo);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> MMA reads
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ah = smem_u32(sA_hi), al = smem_u32(sA_lo), bh = smem_u32(sB_hi), bl = smem_u32(sB_lo);
#pragma unroll
                for (int part = 0; part < 3; ++part) {  // hi.hi, hi.lo, lo.hi
                    const uint32_t pa = part == 2 ? al : ah, pb = part == 1 ? bl : bh;
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k)
                        mma_bf16(tmem, smem_desc(pa + 32 * k), smem_desc(pb + 32 * k), (ch | part | k) != 0);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 smem_u32(s_bar))
                             : "memory");
            }
            ok = mbar_wait_bounded(s_bar, phase);  // MMAs done: operands reusable, accumulators readable
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (!ok) break;
        // epilogue: this thread's row, 32 columns at a time
#pragma unroll 1
        for (int c0 = col_half * (TB / 2); c0 < (col_half + 1) * (TB / 2); c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // lower bound of the exact distance minus ||q||^2: ||x||^2 - 2 dot - err.  All 32 tests
            // first (no side effects: the shared loads batch), the rare appends afterwards.
            uint32_t hits = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float2 xs = s_x[c0 + j];
                float lb = fmaf(-2.0f, __uint_as_float(r[j]), xs.x);
                lb = fmaf(-xs.y, qerr, lb);
                hits |= (lb <= thr ? 1u : 0u) << j;
            }
            while (hits) {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                const uint32_t slot = atomicAdd(&a.cand_cnt[q], 1u);
                if (slot < (uint32_t)a.cand_cap) a.cand[(size_t)q * a.cand_cap + slot] = (uint32_t)(x0 + c0 + j);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // every row is read before the next tile's first MMA overwrites the columns
    }
    if (!ok && threadIdx.x == 0) atomicExch(a.error, 1u);
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TB) : "memory");
}

// One warp per query: exact distance (pmain:150-156) of every candidate, k best keys
// (distance bits << 32 | id) merged into the sorted running state.  A query whose candidate
// list overflowed is left untouched and reported in `flagged` (the dense path redoes it).
constexpr int RS_WARPS = 4, RS_BUF = 160;  // topk <= 64: state + up to 96 pending keys
__global__ void __launch_bounds__(RS_WARPS * 32) gt_rescore_kernel(const GtTcArgs a, int64_t id0, int topk,
                                                                   unsigned long long* __restrict__ state,
                                                                   uint32_t* __restrict__ flagged,
                                                                   uint32_t* __restrict__ n_flagged) {
    __shared__ unsigned long long s_buf[RS_WARPS][RS_BUF];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * RS_WARPS + w;
    if (q >= a.Q) return;
    const uint32_t cnt = a.cand_cnt[q];
    if (cnt > (uint32_t)a.cand_cap) {
        if (lane == 0) flagged[atomicAdd(n_flagged, 1u)] = (uint32_t)q;
        return;
    }
    unsigned long long* buf = s_buf[w];
    unsigned long long* st = state + (size_t)q * topk;
    for (int i = lane; i < topk; i += 32) buf[i] = st[i];
    __syncwarp();
    int nb = topk;
    unsigned long long bound = buf[topk - 1];
    const float* qv = a.queries + (size_t)q * a.D;
    auto compact = [&]() {  // rank counting: keep the topk smallest of buf[0..nb), sorted
        unsigned long long mine[RS_BUF / 32];
        int rank[RS_BUF / 32];
#pragma unroll
        for (int t = 0; t < RS_BUF / 32; ++t) {
            mine[t] = t * 32 + lane < nb ? buf[t * 32 + lane] : ~0ull;
            rank[t] = 0;
        }
        __syncwarp();
        for (int i = 0; i < nb; ++i) {  // equal keys (the initial sentinels) rank by buffer index
            const unsigned long long o = buf[i];
#pragma unroll
            for (int t = 0; t < RS_BUF / 32; ++t) rank[t] += o < mine[t] || (o == mine[t] && i < t * 32 + lane);
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < RS_BUF / 32; ++t)
            if (t * 32 + lane < nb && rank[t] < topk) buf[rank[t]] = mine[t];
        __syncwarp();
        nb = topk;
        bound = buf[topk - 1];
    };
    for (uint32_t i0 = 0; i0 < cnt; i0 += 32) {
        unsigned long long key = ~0ull;
        if (i0 + lane < cnt) {
            const uint32_t v = a.cand[(size_t)q * a.cand_cap + i0 + lane];
            const float* x = a.base + (size_t)v * a.D;
            double acc = 0.0;
            for (int d = 0; d < a.D; ++d) {
                const float diff = __fsub_rn(x[d], qv[d]);
                acc = __dadd_rn(acc, (double)__fmul_rn(diff, diff));
            }
            key = ((unsigned long long)__float_as_uint((float)acc) << 32) | (unsigned long long)(uint32_t)(id0 + v);
        }
        const bool take = key < bound;
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (take) buf[nb + __popc(m & ((1u << lane) - 1u))] = key;
        nb += __popc(m);
        __syncwarp();
        if (nb + 32 > RS_BUF) compact();
    }
    if (nb > topk) compact();
    for (int i = lane; i < topk; i += 32) st[i] = buf[i];
}

// ------------------------------------------------------------------------------------------
// Pipelined form.  The operands are split ONCE per launch into bf16 hi / lo blocks laid out in
// global memory exactly as the MMA wants them in shared memory (gt_split_kernel), so the filter's
// inner loop moves them with TMA bulk copies and converts nothing:
//
//   warp 0 (one lane)  producer: per (tile, chunk) four cp.async.bulk copies (A hi, A lo, B hi,
//                      B lo; 96 KB) into one of two stages, completion on full[stage]
//   warp 1 (one lane)  MMA issuer: waits full[stage], 12 tcgen05.mma into accumulator buffer
//                      tile & 1 (2 x 256 TMEM columns), tcgen05.commit -> empty[stage]; after a
//                      tile's last chunk a second commit -> acc_full[buffer]
//   warps 2..5         epilogue: wait acc_full, read their 32 TMEM lanes (tcgen05.ld), test,
//                      append candidates, arrive on acc_empty[buffer]
// so the copies of chunk c + 1, the MMAs of chunk c and the epilogue of the previous tile overlap.
// Block layout in global memory: queries [query block][chunk][hi, lo][16 KB],
// base [tile][chunk][hi, lo][32 KB].
constexpr int V2_STAGES = 2;
constexpr int V2_STAGE_BYTES = 2 * A_PART + 2 * B_PART;                     // 96 KB
constexpr int V2_SMEM = V2_STAGES * V2_STAGE_BYTES + 2 * TB * 8 + 128;       // + norms of two tiles + barriers
constexpr int V2_THREADS = 192;

template <int ROWS>
__global__ void __launch_bounds__(NT) gt_split_kernel(const float* __restrict__ src, int64_t n_rows, int D,
                                                      int n_chunks, unsigned char* __restrict__ dst) {
    const int64_t tile = blockIdx.x;
    const int ch = blockIdx.y;
    OperandFill<ROWS> f;
    f.load(src, tile * ROWS, n_rows, D, ch * KC);
    unsigned char* blk = dst + ((size_t)tile * n_chunks + ch) * (size_t)(2 * ROWS * KC * 2);
    f.store(blk, blk + ROWS * KC * 2);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(V2_THREADS, 1) gt_tc_filter2_kernel(const GtTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* s_x = reinterpret_cast<float2*>(smem + V2_STAGES * V2_STAGE_BYTES);  // [2][TB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + V2_STAGES * V2_STAGE_BYTES + 2 * TB * 8);
    uint64_t* full = bars;                  // [V2_STAGES]
    uint64_t* empty = bars + V2_STAGES;     // [V2_STAGES]
    uint64_t* acc_full = bars + 2 * V2_STAGES;      // [2]
    uint64_t* acc_empty = bars + 2 * V2_STAGES + 2;  // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * V2_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.y;
    const int64_t n_tiles = (a.n + TB - 1) / TB;
    const int64_t t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;
    const int n_chunks = (a.D + KC - 1) / KC;

    if (threadIdx.x == 0) {
        for (int i = 0; i < V2_STAGES; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(acc_full + i, 1);
            mbar_init(acc_empty + i, 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // both accumulator buffers: all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    bool ok = (smem_u32(smem) & 1023u) == 0;  // the swizzle is a function of the absolute address

    if (warp == 0) {
        if (lane == 0) {  // ---- producer
            uint32_t it = 0;
            for (int64_t t = t_lo; t < t_hi && ok; ++t)
                for (int ch = 0; ch < n_chunks && ok; ++ch, ++it) {
                    const uint32_t stage = it % V2_STAGES, n = it / V2_STAGES;
                    ok = mbar_wait_bounded(empty + stage, (n & 1u) ^ 1u);
                    if (!ok) break;
                    unsigned char* st = smem + stage * V2_STAGE_BYTES;
                    const unsigned char* qa = a.q_split + ((size_t)qb * n_chunks + ch) * (size_t)(2 * A_PART);
                    const unsigned char* xb = a.x_split + ((size_t)t * n_chunks + ch) * (size_t)(2 * B_PART);
                    mbar_expect_tx(full + stage, (uint32_t)V2_STAGE_BYTES);
                    bulk_g2s(st, qa, 2 * A_PART, full + stage);                        // A hi + lo (contiguous)
                    bulk_g2s(st + 2 * A_PART, xb, B_PART, full + stage);               // B hi
                    bulk_g2s(st + 2 * A_PART + B_PART, xb + B_PART, B_PART, full + stage);  // B lo
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            uint32_t it = 0, tc = 0;
            for (int64_t t = t_lo; t < t_hi && ok; ++t, ++tc) {
                const uint32_t buf = tc & 1u, m = tc >> 1;
                ok = mbar_wait_bounded(acc_empty + buf, (m & 1u) ^ 1u);  // the epilogue drained this buffer
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem + buf * (uint32_t)TB;
                for (int ch = 0; ch < n_chunks && ok; ++ch, ++it) {
                    const uint32_t stage = it % V2_STAGES, n = it / V2_STAGES;
                    ok = mbar_wait_bounded(full + stage, n & 1u);
                    if (!ok) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t ah = smem_u32(smem + stage * V2_STAGE_BYTES), al = ah + A_PART;
                    const uint32_t bh = ah + 2 * A_PART, bl = bh + B_PART;
#pragma unroll
                    for (int part = 0; part < 3; ++part) {  // hi.hi, hi.lo, lo.hi
                        const uint32_t pa = part == 2 ? al : ah, pb = part == 1 ? bl : bh;
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)
                            mma_bf16(acc, smem_desc(pa + 32 * k), smem_desc(pb + 32 * k), (ch | part | k) != 0);
                    }
                    umma_commit(empty + stage);  // the stage is free once these MMAs have read it
                }
                if (ok) umma_commit(acc_full + buf);
            }
        }
    } else {  // ---- epilogue warps 2..5: TMEM lane group = warp & 3
        const int lane_grp = warp & 3;
        const int q = qb * TQ + lane_grp * 32 + lane;
        const float thr = q < a.Q ? a.thr[q] : -FLT_MAX;
        const float qerr = q < a.Q ? a.qerr[q] : 0.0f;
        const int et = threadIdx.x - 64;  // 0..127
        uint32_t tc = 0;
        for (int64_t t = t_lo; t < t_hi && ok; ++t, ++tc) {
            const uint32_t buf = tc & 1u, m = tc >> 1;
            const int64_t x0 = t * TB;
            float2* sx = s_x + buf * TB;
            // the buffer's previous tile was fully processed by all four warps two tiles ago (the
            // MMA of this tile could only start after their acc_empty arrivals)
            for (int c = et; c < TB; c += 128)
                sx[c] = x0 + c < a.n ? make_float2(a.x_nlo[x0 + c], a.x_len[x0 + c])
                                     : make_float2(__int_as_float(0x7f800000), 0.0f);
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
            ok = mbar_wait_bounded(acc_full + buf, m & 1u);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = 0; c0 < TB; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(lane_grp * 32) << 16) + buf * (uint32_t)TB + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                      "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                      "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                      "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t hits = 0;  // all 32 tests first (the shared loads batch), the rare appends afterwards
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float2 xs = sx[c0 + j];
                    float lb = fmaf(-2.0f, __uint_as_float(r[j]), xs.x);
                    lb = fmaf(-xs.y, qerr, lb);
                    hits |= (lb <= thr ? 1u : 0u) << j;
                }
                while (hits) {
                    const int j = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const uint32_t slot = atomicAdd(&a.cand_cnt[q], 1u);
                    if (slot < (uint32_t)a.cand_cap) a.cand[(size_t)q * a.cand_cap + slot] = (uint32_t)(x0 + c0 + j);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);
        }
    }
    if (!ok) atomicExch(a.error, 1u);
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

cudaError_t launch_gt_split(const float* src, int64_t n_rows, int D, bool queries, unsigned char* dst, cudaStream_t st) {
    const int n_chunks = (D + KC - 1) / KC;
    if (queries) {
        const int64_t blocks = (n_rows + TQ - 1) / TQ;
        gt_split_kernel<TQ><<<dim3((unsigned)blocks, (unsigned)n_chunks), NT, 0, st>>>(src, n_rows, D, n_chunks, dst);
    } else {
        const int64_t tiles = (n_rows + TB - 1) / TB;
        gt_split_kernel<TB><<<dim3((unsigned)tiles, (unsigned)n_chunks), NT, 0, st>>>(src, n_rows, D, n_chunks, dst);
    }
    return cudaGetLastError();
}

size_t gt_split_bytes(int64_t n_rows, int D, bool queries) {
    const size_t n_chunks = (size_t)(D + KC - 1) / KC;
    const size_t rows = queries ? TQ : TB;
    return (size_t)((n_rows + (int64_t)rows - 1) / (int64_t)rows) * n_chunks * 2 * rows * KC * 2;
}

cudaError_t launch_gt_tc_filter2(const GtTcArgs& a, int n_sms, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(gt_tc_filter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM);
    if (e != cudaSuccess) return e;
    const int qb = (a.Q + TQ - 1) / TQ;
    const int64_t n_tiles = (a.n + TB - 1) / TB;
    // one persistent CTA per SM: split the base range so the grid is close to whole waves
    const int64_t slots = n_sms;
    int64_t splits = 1;
    double best = -1.0;
    for (int w = 1; w <= 8; ++w) {
        int64_t sp = slots * w / qb;
        if (sp < 1) sp = 1;
        if (sp > n_tiles) sp = n_tiles;
        const int64_t ctas = sp * qb, waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots);
        if (eff > best + 1e-9) {
            best = eff;
            splits = sp;
        }
    }
    gt_tc_filter2_kernel<<<dim3((unsigned)splits, (unsigned)qb), V2_THREADS, V2_SMEM, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_gt_prep(const float* x, int64_t n, int D, float* nlo, float* nhi, float* len, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    gt_prep_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, n, D, nlo, nhi, len);
    return cudaGetLastError();
}

cudaError_t launch_gt_thr(const unsigned long long* state, int topk, int Q, const float* q_nlo, const float* q_nhi,
                          const float* q_len, float c_err, float* thr, float* qerr, cudaStream_t st) {
    gt_thr_kernel<<<(Q + 127) / 128, 128, 0, st>>>(state, topk, Q, q_nlo, q_nhi, q_len, c_err, thr, qerr);
    return cudaGetLastError();
}

cudaError_t launch_gt_tc_filter(const GtTcArgs& a, int n_sms, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(gt_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    const int qb = (a.Q + TQ - 1) / TQ;
    const int64_t n_tiles = (a.n + TB - 1) / TB;
    // two CTAs per SM (96 KB of operands, 256 TMEM columns each): split the base range so that the
    // grid is as close as possible to a whole number of waves of 2 * n_sms CTAs
    const int64_t slots = 2LL * n_sms;
    int64_t splits = 1;
    double best = -1.0;
    for (int w = 2; w <= 8; ++w) {
        int64_t sp = slots * w / qb;
        if (sp < 1) sp = 1;
        if (sp > n_tiles) sp = n_tiles;
        const int64_t ctas = sp * qb, waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots);
        if (eff > best + 1e-9) {
            best = eff;
            splits = sp;
        }
    }
    gt_tc_filter_kernel<<<dim3((unsigned)splits, (unsigned)qb), NT, SMEM_TOTAL, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_gt_rescore(const GtTcArgs& a, int64_t id0, int topk, unsigned long long* state, uint32_t* flagged,
                              uint32_t* n_flagged, cudaStream_t st) {
    gt_rescore_kernel<<<(a.Q + RS_WARPS - 1) / RS_WARPS, RS_WARPS * 32, 0, st>>>(a, id0, topk, state, flagged, n_flagged);
    return cudaGetLastError();
}

}  // namespace dpq
