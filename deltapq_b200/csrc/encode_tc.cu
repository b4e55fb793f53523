// PQ encode on the tensor cores (reference PQTree::EncodePlain, pq_tree.cpp:192-253).
//
// The reference takes, per (vector, subspace), the argmin over K centroids of a float distance
// accumulated dimension by dimension (float subtract, float multiply, float add, strict <).  That is
// 3 K Ds FP32 operations per code, FP32-issue bound on the SIMT encoder (secondary.cu).  Here the K
// scores of 128 vectors come from ONE tcgen05.mma group per subspace:
//
//      ||x - c_k||^2 = ||x||^2 - 2 (x.c_k - ||c_k||^2 / 2)       =>   argmin_k d = argmax_k v_k
//
// A' row (vector)   = [ x_hi | x_hi | x_lo | 1 1 1 | 0.. ]   (bf16, 64 columns = one 128-byte swizzle row)
// B' row (centroid) = [ c_hi | c_lo | c_hi | n1 n2 n3 | 0.. ],  n1 + n2 + n3 = -||c||^2 / 2 split three ways
// so that A'.B' = x_hi.c_hi + x_hi.c_lo + x_lo.c_hi - ||c||^2 / 2 = v_k up to a rigorously bounded error E.
// The MMA result is a FILTER: every centroid whose score is within the bound of the best score (in
// practice: every group of four consecutive centroids whose best score is) is re-scored in the
// reference's own arithmetic, in ascending centroid order with strict <, so the code written is the
// reference's, bit for bit, whatever the tensor cores rounded (tests compare
// with the SIMT kernel and the oracle on adversarial inputs: duplicated centroids, exact ties,
// huge / tiny magnitudes, non-finite values).
//
// Error bound (u = 2^-24; bf16 keeps 8 significant bits):
//   split:   x = x_hi + x_lo + r_x, |x - x_hi| <= 2^-9 |x|, |r_x| <= 2^-18 |x| (same for c); the three
//            dropped products x_lo.c_lo, r_x.c, x.r_c are <= 3.02 * 2^-18 |x||c| per dimension, and
//            sum_d |x_d||c_d| <= ||x|| ||c||;
//   norm:    the three-way split of -||c||^2 / 2 (computed in double) leaves <= 2^-26 ||c||^2 / 2;
//   fp32 accumulation of <= 64 exact products inside the tensor core: <= (64 + 8) * 2^-23 of the sum
//            of magnitudes (the same allowance the ground-truth filter uses, gt_tc.cu / secondary.cu);
//   =>  |v_k - (x.c_k - ||c_k||^2/2)| <= E = 2.3e-5 * (||x|| cmax + cmax^2 / 2),  cmax = max_k ||c_k||.
//   reference arithmetic: d_ref = d_true (1 +- g), g = (Ds + 3) * 2^-24 * 1.001 (non-negative terms).
//   The reference's winner k* has d_ref(k*) <= d_ref(kb) for the filter's best column kb, hence
//   d_true(k*) <= d_true(kb) (1 + 3 g)  =>  v_k* >= v_kb - 2 E - 1.5 g d_true(kb),
//   and d_true(kb) <= ||x||^2 - 2 v_kb + 2 E.  Rows where any of this is not finite, or so small that
//   the bound underflows, take every column as a candidate (= the reference loop itself).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "encode_tc.cuh"
#include "umma.cuh"

namespace dpq {
namespace {

constexpr int ET_ROWS = 128;  // vectors per tile = MMA M = TMEM lanes
constexpr int ET_N = 256;     // centroid columns = MMA N = TMEM columns
constexpr int ET_A_BYTES = ET_ROWS * 128;  // 16 KB
constexpr int ET_B_BYTES = ET_N * 128;     // 32 KB
constexpr uint32_t ET_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ET_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr int ET_THREADS = 256;  // two threads per row: column halves
constexpr int ET_SMEM_MIN = 80 * 1024;  // > 228 KB / 3: at most two CTAs per SM, 256 TMEM columns each

// bounded spin on the MMA completion barrier (a few seconds at most: a failure is reported, never a hang)
__device__ __forceinline__ bool wait_mma(uint64_t* bar, uint32_t phase) {
    for (uint32_t spin = 0; spin < (1u << 21); ++spin) {
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

__device__ __forceinline__ uint16_t bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_val(uint16_t b) { return __bfloat162float(__ushort_as_bfloat16(b)); }

// element e of an operand row.  A: [hi | hi | lo | 1 1 1 | 0..], B: [hi | lo | hi | tail | 0..]
template <int DS, bool IS_B>
__device__ __forceinline__ uint16_t row_elem(int e, const uint16_t (&hi)[DS], const uint16_t (&lo)[DS], const uint16_t (&tail)[3]) {
    if (e < DS) return hi[e];
    if (e < 2 * DS) return IS_B ? lo[e - DS] : hi[e - DS];
    if (e < 3 * DS) return IS_B ? hi[e - 2 * DS] : lo[e - 2 * DS];
    if (e < 3 * DS + 3) return tail[e - 3 * DS];
    return 0;
}

template <int DS, bool IS_B>
__device__ __forceinline__ void store_row(unsigned char* tile, int row, const uint16_t (&hi)[DS], const uint16_t (&lo)[DS],
                                          const uint16_t (&tail)[3]) {
    constexpr int PIECES = ((3 * DS + 3 + 15) / 16) * 2;  // 16-byte pieces the MMA k-steps read
#pragma unroll
    for (int p = 0; p < PIECES; ++p) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            w[i] = (uint32_t)row_elem<DS, IS_B>(p * 8 + 2 * i, hi, lo, tail) |
                   ((uint32_t)row_elem<DS, IS_B>(p * 8 + 2 * i + 1, hi, lo, tail) << 16);
        *reinterpret_cast<uint4*>(tile + umma::swz_off(row, p)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// B' blocks [M][256 rows x 128 bytes] in the shared-memory layout, and cmax[m] >= max_k ||c_k||
template <int DS>
__global__ void __launch_bounds__(ET_N) encode_tc_prep_kernel(const float* __restrict__ cw, int K, unsigned char* __restrict__ bsplit,
                                                              float* __restrict__ cmax) {
    __shared__ float s_len[ET_N / 32];
    const int m = blockIdx.x, k = threadIdx.x;
    uint16_t hi[DS], lo[DS], tail[3] = {0, 0, 0};
    double n2 = 0.0;
#pragma unroll
    for (int d = 0; d < DS; ++d) {
        const float c = k < K ? cw[((size_t)m * K + k) * DS + d] : 0.0f;
        hi[d] = bf16_bits(c);
        lo[d] = bf16_bits(c - bf16_val(hi[d]));
        n2 += (double)c * (double)c;
    }
    double rest = -0.5 * n2;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        tail[j] = bf16_bits((float)rest);
        rest -= (double)bf16_val(tail[j]);
    }
    store_row<DS, true>(bsplit + (size_t)m * ET_B_BYTES, k, hi, lo, tail);
    float len = __double2float_ru(sqrt(n2) * (1.0 + 1e-6));  // NaN / inf propagate: the main kernel then takes every column
    for (int o = 16; o; o >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, len, o);
        len = (len != len || other != other) ? __int_as_float(0x7fc00000) : fmaxf(len, other);
    }
    if ((k & 31) == 0) s_len[k >> 5] = len;
    __syncthreads();
    if (k == 0) {
        float v = s_len[0];
        for (int w = 1; w < ET_N / 32; ++w) v = (v != v || s_len[w] != s_len[w]) ? __int_as_float(0x7fc00000) : fmaxf(v, s_len[w]);
        cmax[m] = v;
    }
}

// the reference's distance, op for op (pq_tree.cpp:215-237): never contracted into FMAs
template <int DS>
__device__ __forceinline__ float ref_dist(const float (&x)[DS], const float* __restrict__ c) {
    float dist = 0.0f;
#pragma unroll
    for (int d = 0; d < DS; d += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(c + d);  // rows are 16-byte aligned
        const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float diff = __fsub_rn(x[d + i], cv[i]);
            dist = __fadd_rn(dist, __fmul_rn(diff, diff));
        }
    }
    return dist;
}

// rows of tile `t` (this CTA's subspace) -> s_x[row][XS] (zero beyond n and beyond D: pq_tree.cpp:194-198).
// Aligned input: 16-byte cp.async copies, completed with cp.async.wait_group by the caller, so the next
// tile's rows travel while this tile is scored.  Thread (r, h) moves the chunks c of row r with c % 2 == h.
template <int DS>
__device__ __forceinline__ void stage_rows(const EncTcArgs& a, int64_t t, int m, float* s_x, int r, int h) {
    constexpr int XS = DS + 4;
    const int64_t row = t * ET_ROWS + r;
    if (a.vec_ok && (m + 1) * DS <= a.D) {
        const bool valid = row < a.n;
        const float* src = a.x + (valid ? (size_t)row * a.D + (size_t)m * DS : 0);
#pragma unroll
        for (int c = 0; c < DS / 4; ++c)
            if ((c & 1) == h)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(s_x + r * XS + 4 * c)), "l"(src + 4 * c),
                             "r"(valid ? 16 : 0)
                             : "memory");
    } else {
#pragma unroll
        for (int c = 0; c < DS / 4; ++c)
            if ((c & 1) == h)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int d = 4 * c + i;
                    s_x[r * XS + d] = (row < a.n && m * DS + d < a.D) ? __ldg(a.x + (size_t)row * a.D + (size_t)m * DS + d) : 0.0f;
                }
    }
}

// CTA = (range of 128-vector tiles, subspace m), 256 threads: thread t owns row r = t & 127 (= TMEM lane)
// and the column half h = t >> 7 (128 centroids), so that an SM holds 16 warps (two CTAs of 256 TMEM
// columns): the first version (128 threads, 8 warps per SM) issued 44 % of its cycles (ncu, stall "wait").
// Per tile: A' from the staged rows -> 1..4 tcgen05.mma (one elected thread) -> every thread reads its 128
// scores ONCE (four tcgen05.ld), keeping the maximum of each group of four columns (32 registers, a
// by-product of the max tree) -> the halves exchange their row maxima -> groups whose maximum reaches the
// candidate limit are re-scored in the reference's arithmetic, four columns each, ascending -> the halves
// exchange (distance, id).  A second pass over TMEM for the individual columns cost more than re-scoring
// three extra columns per candidate (tcgen05.ld is warp-collective: a warp visits the union of its rows'
// candidate blocks, i.e. nearly all of them).
template <int DS, bool FULLK>  // FULLK: K == 256, no column masks
__global__ void __launch_bounds__(ET_THREADS, 2) encode_tc_kernel(const EncTcArgs a) {
    constexpr int KSTEPS = (3 * DS + 3 + 15) / 16;
    constexpr int PIECES = 2 * KSTEPS;
    constexpr int XS = DS + 4;  // row stride of the staged rows: 16-byte accesses of 8 consecutive rows hit 8 bank groups
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + ET_B_BYTES;
    float* s_cw = reinterpret_cast<float*>(smem + ET_B_BYTES + ET_A_BYTES);  // [K][XS] exact centroids, rows padded like s_x:
    // the rows of a warp's 32 different candidates then spread over the banks (ncu, unpadded: 60 % of the
    // kernel's shared-memory wavefronts were bank conflicts, top stall short_scoreboard)
    float* s_xb = s_cw + ET_N * XS;                                           // [2][ET_ROWS][XS] staged rows
    float* s_vmax = s_xb + 2 * ET_ROWS * XS;                                  // [256] row maxima per half
    float* s_best = s_vmax + ET_THREADS;                                      // [128] upper half's best distance
    int* s_bestk = reinterpret_cast<int*>(s_best + ET_ROWS);                  // [128] and its centroid
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bestk + ET_ROWS);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);

    const int tid = threadIdx.x, r = tid & (ET_ROWS - 1), h = tid >> 7, warp = tid >> 5, m = blockIdx.y, K = a.K;
    const int64_t n_tiles = (a.n + ET_ROWS - 1) / ET_ROWS;
    const int64_t t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;

    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)ET_N)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {  // the subspace's B' block (already in the shared-memory layout) and its exact centroids
        const uint4* src = reinterpret_cast<const uint4*>(a.bsplit + (size_t)m * ET_B_BYTES);
        uint4* dst = reinterpret_cast<uint4*>(sB);
        for (int i = tid; i < ET_B_BYTES / 16; i += ET_THREADS) dst[i] = __ldg(src + i);
        for (int i = tid; i < K * DS; i += ET_THREADS) s_cw[(i / DS) * XS + i % DS] = a.cw[(size_t)m * K * DS + i];
    }
    if (t_lo < t_hi) stage_rows<DS>(a, t_lo, m, s_xb, r, h);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * (ET_N / 2));

    const float cmax = a.cmax[m];
    const float gam = (float)(DS + 3) * 5.97e-8f;  // (Ds + 3) * 2^-24 * 1.001
    uint32_t phase = 0;
    bool ok = (smem_u32(smem) & 1023u) == 0;  // the swizzle is a function of the absolute address

    for (int64_t t = t_lo; t < t_hi && ok; ++t) {
        const float* s_x = s_xb + (int)((t - t_lo) & 1) * ET_ROWS * XS + r * XS;
        float* s_xnext = s_xb + (int)((t - t_lo + 1) & 1) * ET_ROWS * XS;
        float xn2 = 0.0f;
        {
            float x[DS];
#pragma unroll
            for (int d = 0; d < DS; d += 4) {
                const float4 v = *reinterpret_cast<const float4*>(s_x + d);
                x[d] = v.x, x[d + 1] = v.y, x[d + 2] = v.z, x[d + 3] = v.w;
            }
#pragma unroll
            for (int d = 0; d < DS; ++d) xn2 = fmaf(x[d], x[d], xn2);
            if constexpr (DS == 16) {
                // the two halves write alternate 16-byte pieces of the row: pieces h, 2 + h (hi) and 4 + h (lo)
                // hold dimensions 8h .. 8h + 7, so each half converts only its eight values
                const float4 u0 = *reinterpret_cast<const float4*>(s_x + 8 * h), u1 = *reinterpret_cast<const float4*>(s_x + 8 * h + 4);
                const float u[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint16_t h0 = bf16_bits(u[2 * i]), h1 = bf16_bits(u[2 * i + 1]);
                    const uint16_t l0 = bf16_bits(u[2 * i] - bf16_val(h0)), l1 = bf16_bits(u[2 * i + 1] - bf16_val(h1));
                    wh[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                    wl[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                }
                const uint4 ph = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                *reinterpret_cast<uint4*>(sA + umma::swz_off(r, h)) = ph;
                *reinterpret_cast<uint4*>(sA + umma::swz_off(r, 2 + h)) = ph;
                *reinterpret_cast<uint4*>(sA + umma::swz_off(r, 4 + h)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                *reinterpret_cast<uint4*>(sA + umma::swz_off(r, 6 + h)) =
                    h == 0 ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);  // 1 1 1 0 ...
            } else {
                uint16_t hi[DS], lo[DS];
                const uint16_t one[3] = {0x3F80, 0x3F80, 0x3F80};
#pragma unroll
                for (int d = 0; d < DS; ++d) {
                    hi[d] = bf16_bits(x[d]);
                    lo[d] = bf16_bits(x[d] - bf16_val(hi[d]));
                }
#pragma unroll
                for (int p = 0; p < PIECES; ++p) {
                    if ((p & 1) == h) {
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            w[i] = (uint32_t)row_elem<DS, false>(p * 8 + 2 * i, hi, lo, one) |
                                   ((uint32_t)row_elem<DS, false>(p * 8 + 2 * i + 1, hi, lo, one) << 16);
                        *reinterpret_cast<uint4*>(sA + umma::swz_off(r, p)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> MMA reads
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(sA), sb = smem_u32(sB);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
                umma::mma_bf16<ET_IDESC>(tmem, umma::smem_desc(sa + 32 * k), umma::smem_desc(sb + 32 * k), k != 0);
            umma::commit(s_bar);
        }
        if (t + 1 < t_hi) stage_rows<DS>(a, t + 1, m, s_xnext, r, h);  // the next tile's rows travel meanwhile
        asm volatile("cp.async.commit_group;" ::: "memory");
        ok = wait_mma(s_bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) break;

        // this thread's 128 scores, once: the maximum of every group of four columns, and of the half row
        float gm[ET_N / 8];
        float vmax = -FLT_MAX;
#pragma unroll
        for (int b = 0; b < ET_N / 64; ++b) {
            uint32_t v[32];
            umma::tmem_ld32(tlane + (uint32_t)(b * 32), v);
            umma::tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float g4 = -FLT_MAX;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = h * (ET_N / 2) + b * 32 + 4 * g + i;  // padded columns (k >= K) hold 0: masked
                    g4 = fmaxf(g4, FULLK || k < K ? __uint_as_float(v[4 * g + i]) : -FLT_MAX);
                }
                gm[b * 8 + g] = g4;
                vmax = fmaxf(vmax, g4);
            }
        }
        s_vmax[tid] = vmax;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // TMEM and sA are free from here on
        vmax = fmaxf(s_vmax[r], s_vmax[r + ET_ROWS]);

        // candidate limit (see the header): everything not provably worse than the best column
        const float xlen = sqrtf(xn2) * 1.0001f, x2 = xn2 * 1.0001f;
        const float scale = fmaf(xlen, cmax, 0.5f * cmax * cmax);
        const float E = 2.3e-5f * scale;
        // upper bound of d_true(kb), the float evaluation's own rounding included
        const float dkb = fmaxf(fmaf(-2.0f, vmax, x2) + 2.0f * E, 0.0f) + 1e-6f * (x2 + 2.0f * fabsf(vmax));
        const float thr = 1.01f * (2.0f * E + 1.5f * gam * dkb);
        const float lim = vmax - thr;
        // not finite, or small enough for operands / products to be flushed or the bound to underflow:
        // every column is a candidate (the loop below is then the reference's loop)
        const bool every = !(scale < 1e30f) || !(scale > 1e-25f) || !(cmax > 1e-25f) || !(x2 < 1e30f) || !(thr < 1e30f) ||
                           !(lim == lim);
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < ET_N / 8; ++j) mask |= (!(gm[j] < lim) ? 1u : 0u) << j;
        if (every) mask = 0xFFFFFFFFu;

        // exact re-score of the candidate groups, ascending centroid id, strict <  (pq_tree.cpp:225-233).  Every
        // row walks ITS groups, so a warp runs as many rounds as its row with the most groups (one, mostly).
        float best = FLT_MAX;
        int best_k = -1;  // nothing below FLT_MAX (overflow, NaN): the reference writes (uchar)-1 (pq_tree.cpp:217, 235)
        if (mask) {
            float x[DS];
#pragma unroll
            for (int d = 0; d < DS; d += 4) {
                const float4 v = *reinterpret_cast<const float4*>(s_x + d);
                x[d] = v.x, x[d + 1] = v.y, x[d + 2] = v.z, x[d + 3] = v.w;
            }
            while (mask) {
                const int g = __ffs(mask) - 1;
                mask &= mask - 1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = h * (ET_N / 2) + 4 * g + i;
                    if (FULLK || k < K) {
                        const float dist = ref_dist<DS>(x, s_cw + k * XS);
                        if (dist < best) {
                            best = dist;
                            best_k = k;
                        }
                    }
                }
            }
        }
        if (h == 1) {
            s_best[r] = best;
            s_bestk[r] = best_k;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // the next tile's rows have landed (this thread's copies)
        __syncthreads();                                      // ... everyone's; and the upper half's results
        if (h == 0) {
            // the lower half holds the smaller ids: the upper half wins only with a strictly smaller distance
            if (s_best[r] < best) best_k = s_bestk[r];
            const int64_t row = t * ET_ROWS + r;
            if (row < a.n) a.codes[(size_t)row * a.M + m] = (uint8_t)best_k;
        }
        // s_best / s_bestk are rewritten only after the next tile's two barriers
    }
    if (!ok && tid == 0) atomicExch(a.error, 1u);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)ET_N) : "memory");
}

template <int DS>
cudaError_t launch_ds(const EncTcArgs& a, int n_sms, cudaStream_t st) {
    encode_tc_prep_kernel<DS><<<a.M, ET_N, 0, st>>>(a.cw, a.K, const_cast<unsigned char*>(a.bsplit), const_cast<float*>(a.cmax));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int smem = ET_B_BYTES + ET_A_BYTES + ET_N * (DS + 4) * 4 + 2 * ET_ROWS * (DS + 4) * 4 + ET_THREADS * 4 + ET_ROWS * 8 + 64;
    if (smem < ET_SMEM_MIN) smem = ET_SMEM_MIN;
    auto kernel = a.K == ET_N ? encode_tc_kernel<DS, true> : encode_tc_kernel<DS, false>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int64_t n_tiles = (a.n + ET_ROWS - 1) / ET_ROWS;
    int64_t gx = (2LL * n_sms + a.M - 1) / a.M;  // two CTAs per SM
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    kernel<<<dim3((unsigned)gx, (unsigned)a.M), ET_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

bool encode_tc_supported(int M, int K, int Ds) { return (Ds == 4 || Ds == 8 || Ds == 16) && K >= 1 && K <= ET_N && M >= 1 && M <= 64; }
size_t encode_tc_scratch_bytes(int M) { return (size_t)M * ET_B_BYTES + (size_t)M * sizeof(float) + 16; }

cudaError_t launch_encode_tc(const float* d_cw, int M, int K, int Ds, const float* d_x, int64_t n, int D, uint8_t* d_codes,
                             unsigned char* d_scratch, uint32_t* d_error, int n_sms, cudaStream_t st) {
    EncTcArgs a;
    a.x = d_x, a.n = n, a.D = D, a.M = M, a.K = K;
    a.cw = d_cw;
    a.bsplit = d_scratch;
    a.cmax = reinterpret_cast<const float*>(d_scratch + (size_t)M * ET_B_BYTES);
    a.codes = d_codes;
    a.error = d_error;
    a.vec_ok = (reinterpret_cast<uintptr_t>(d_x) & 15u) == 0 && (D & 3) == 0 && (Ds & 3) == 0;
    switch (Ds) {
        case 4: return launch_ds<4>(a, n_sms, st);
        case 8: return launch_ds<8>(a, n_sms, st);
        case 16: return launch_ds<16>(a, n_sms, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dpq
