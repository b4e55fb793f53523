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
// The MMA result is a FILTER: every centroid whose score is within the bound of the best score is
// re-scored in the reference's own arithmetic, in ascending centroid order with strict <, so the
// code written is the reference's, bit for bit, whatever the tensor cores rounded (tests compare
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
constexpr int ET_QCAP = 16;   // candidate queue per row (bytes of shared memory)
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

template <int DS>
__device__ __forceinline__ void load_row(const EncTcArgs& a, int64_t row, int m, float (&x)[DS]) {
    if (row >= a.n) {
#pragma unroll
        for (int d = 0; d < DS; ++d) x[d] = 0.0f;
        return;
    }
    const float* p = a.x + (size_t)row * a.D + (size_t)m * DS;
    if (a.vec_ok && (m + 1) * DS <= a.D) {
#pragma unroll
        for (int d = 0; d < DS; d += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + d));
            x[d] = v.x, x[d + 1] = v.y, x[d + 2] = v.z, x[d + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int d = 0; d < DS; ++d) x[d] = m * DS + d < a.D ? __ldg(p + d) : 0.0f;  // zero padding (pq_tree.cpp:194-198)
    }
}

// the reference's distance, op for op (pq_tree.cpp:215-237): never contracted into FMAs
template <int DS>
__device__ __forceinline__ float ref_dist(const float (&x)[DS], const float* __restrict__ c) {
    float dist = 0.0f;
#pragma unroll
    for (int d = 0; d < DS; ++d) {
        const float diff = __fsub_rn(x[d], c[d]);
        dist = __fadd_rn(dist, __fmul_rn(diff, diff));
    }
    return dist;
}

// CTA = (range of 128-vector tiles, subspace m); thread r owns row r of the tile = TMEM lane r.
template <int DS, bool FULLK>  // FULLK: K == 256, no column masks
__global__ void __launch_bounds__(ET_ROWS) encode_tc_kernel(const EncTcArgs a) {
    constexpr int KSTEPS = (3 * DS + 3 + 15) / 16;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + ET_B_BYTES;
    float* s_cw = reinterpret_cast<float*>(smem + ET_B_BYTES + ET_A_BYTES);  // [K][DS] exact centroids
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + ET_B_BYTES + ET_A_BYTES + ET_N * DS * 4);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);
    uint8_t* s_queue = reinterpret_cast<uint8_t*>(s_bar + 2);  // [ET_ROWS][ET_QCAP]

    const int r = threadIdx.x, warp = r >> 5, m = blockIdx.y, K = a.K;
    const int64_t n_tiles = (a.n + ET_ROWS - 1) / ET_ROWS;
    const int64_t t_lo = n_tiles * blockIdx.x / gridDim.x, t_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;

    if (r == 0) {
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
        for (int i = r; i < ET_B_BYTES / 16; i += ET_ROWS) dst[i] = __ldg(src + i);
        for (int i = r; i < K * DS; i += ET_ROWS) s_cw[i] = a.cw[(size_t)m * K * DS + i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);

    const float cmax = a.cmax[m];
    const float gam = (float)(DS + 3) * 5.97e-8f;  // (Ds + 3) * 2^-24 * 1.001
    uint32_t phase = 0;
    bool ok = (smem_u32(smem) & 1023u) == 0;  // the swizzle is a function of the absolute address

    float x[DS];
    if (t_lo < t_hi) load_row<DS>(a, t_lo * ET_ROWS + r, m, x);
    for (int64_t t = t_lo; t < t_hi && ok; ++t) {
        const int64_t row = t * ET_ROWS + r;
        float xn2 = 0.0f;
        {
            uint16_t hi[DS], lo[DS];
            const uint16_t one[3] = {0x3F80, 0x3F80, 0x3F80};
#pragma unroll
            for (int d = 0; d < DS; ++d) {
                hi[d] = bf16_bits(x[d]);
                lo[d] = bf16_bits(x[d] - bf16_val(hi[d]));
                xn2 = fmaf(x[d], x[d], xn2);
            }
            store_row<DS, false>(sA, r, hi, lo, one);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> MMA reads
        __syncthreads();
        if (r == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(sA), sb = smem_u32(sB);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
                umma::mma_bf16<ET_IDESC>(tmem, umma::smem_desc(sa + 32 * k), umma::smem_desc(sb + 32 * k), k != 0);
            umma::commit(s_bar);
        }
        float xnext[DS];  // the next tile's row travels while the MMAs run
        if (t + 1 < t_hi) load_row<DS>(a, row + ET_ROWS, m, xnext);
        ok = wait_mma(s_bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) break;

        // pass 1: best score of the row, and of each block of 32 columns
        float bm[ET_N / 32];
        float vmax = -FLT_MAX;
#pragma unroll
        for (int b = 0; b < ET_N / 32; ++b) {
            bm[b] = -FLT_MAX;
            if (FULLK || b * 32 < K) {
                uint32_t v[32];
                umma::tmem_ld32(tlane + (uint32_t)(b * 32), v);
                umma::tmem_ld_wait();
                float m0 = -FLT_MAX, m1 = -FLT_MAX, m2 = -FLT_MAX, m3 = -FLT_MAX;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    m0 = fmaxf(m0, FULLK || b * 32 + j < K ? __uint_as_float(v[j]) : -FLT_MAX);
                    m1 = fmaxf(m1, FULLK || b * 32 + j + 1 < K ? __uint_as_float(v[j + 1]) : -FLT_MAX);
                    m2 = fmaxf(m2, FULLK || b * 32 + j + 2 < K ? __uint_as_float(v[j + 2]) : -FLT_MAX);
                    m3 = fmaxf(m3, FULLK || b * 32 + j + 3 < K ? __uint_as_float(v[j + 3]) : -FLT_MAX);
                }
                bm[b] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                vmax = fmaxf(vmax, bm[b]);
            }
        }
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

        // pass 2: the row's candidate columns, ascending, into its private queue.  Groups of four columns are
        // tested by their maximum first: a row has one to three candidates among 256 columns, so nearly every
        // group is dismissed with one compare.  Rows whose bound cannot be trusted skip this and take every column.
        int nc = every ? ET_QCAP + 1 : 0;
        uint8_t* queue = s_queue + r * ET_QCAP;
#pragma unroll
        for (int b = 0; b < ET_N / 32; ++b) {
            // tcgen05.ld is warp-collective (.sync.aligned): the skip must be decided by the whole warp.  A row
            // without a candidate in this block dismisses its eight groups below.
            if ((FULLK || b * 32 < K) && __any_sync(0xffffffffu, !every && !(bm[b] < lim))) {
                uint32_t v[32];
                umma::tmem_ld32(tlane + (uint32_t)(b * 32), v);
                umma::tmem_ld_wait();
                if (!every) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float g4 = fmaxf(fmaxf(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1])),
                                               fmaxf(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3])));
                        if (!(g4 < lim)) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int k = b * 32 + 4 * g + i;
                                if ((FULLK || k < K) && !(__uint_as_float(v[4 * g + i]) < lim)) {
                                    if (nc < ET_QCAP) queue[nc] = (uint8_t)k;
                                    ++nc;
                                }
                            }
                        }
                    }
                }
                __syncwarp();  // the rows' work differs: reconverge before the next collective load
            }
        }
        // exact re-score, ascending centroid id, strict <  (pq_tree.cpp:225-233).  Every row walks ITS queue, so
        // a warp runs as many rounds as its longest queue; a row with more candidates than the queue holds
        // (duplicated centroids en masse, untrusted bound) runs the reference loop over all K columns.
        float best = FLT_MAX;
        int best_k = -1;  // nothing below FLT_MAX (overflow, NaN): the reference writes (uchar)-1 (pq_tree.cpp:217, 235)
        if (nc > ET_QCAP) {
            for (int k = 0; k < K; ++k) {
                const float dist = ref_dist<DS>(x, s_cw + k * DS);
                if (dist < best) {
                    best = dist;
                    best_k = k;
                }
            }
        } else {
            for (int i = 0; i < nc; ++i) {
                const int k = queue[i];
                const float dist = ref_dist<DS>(x, s_cw + k * DS);
                if (dist < best) {
                    best = dist;
                    best_k = k;
                }
            }
        }
        __syncwarp();
        if (row < a.n) a.codes[(size_t)row * a.M + m] = (uint8_t)best_k;
#pragma unroll
        for (int d = 0; d < DS; ++d) x[d] = xnext[d];
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // every row is read before the next tile's MMA overwrites the columns (and sA)
    }
    if (!ok && r == 0) atomicExch(a.error, 1u);
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)ET_N) : "memory");
}

template <int DS>
cudaError_t launch_ds(const EncTcArgs& a, int n_sms, cudaStream_t st) {
    encode_tc_prep_kernel<DS><<<a.M, ET_N, 0, st>>>(a.cw, a.K, const_cast<unsigned char*>(a.bsplit), const_cast<float*>(a.cmax));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int smem = ET_B_BYTES + ET_A_BYTES + ET_N * DS * 4 + 16 + ET_ROWS * ET_QCAP;
    if (smem < ET_SMEM_MIN) smem = ET_SMEM_MIN;
    auto kernel = a.K == ET_N ? encode_tc_kernel<DS, true> : encode_tc_kernel<DS, false>;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int64_t n_tiles = (a.n + ET_ROWS - 1) / ET_ROWS;
    int64_t gx = (2LL * n_sms + a.M - 1) / a.M;  // two CTAs per SM
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    kernel<<<dim3((unsigned)gx, (unsigned)a.M), ET_ROWS, smem, st>>>(a);
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
