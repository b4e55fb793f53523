// Coarse search: the whole tree is scanned with a small-integer table in packed 8-bit arithmetic
// -- four queries per 32-bit word, sixteen per 128-bit table read.  Narrow shape (M <= 8): 5-bit
// entries, 112 queries per CTA, i.e. HALF the shared-memory wavefronts per (node, query) of the
// 15-bit scan (scan2.cu), which is bound by exactly those wavefronts (profiles/r1_summary.md).
// Wide shape (M <= 16): 4-bit entries (sixteen of them fit a byte sum), 48 queries per CTA against
// the 15-bit wide scan's 24; its filter is weaker (saturation at 15 units, 16 x 0.5 rounding slack),
// so it is chosen automatically only for short result lists (api.cu).
//
// Why it is exact.  Before this pass the 15-bit scan runs over a 1/16 sample of the tree and
// select_kernel / the exact fallback give cap_q = the exact k-th distance over the sample, an
// upper bound of the true k-th distance T_q.  Coarse entry = min(31, rint(lut / unit)) with
// unit = cap_q / L (L = 80 levels; entries saturate at 31 so that eight of them never carry out
// of a byte).  Saturation only lowers a sum, and rounding adds at most 0.5 per entry, so a node
// with distance d <= T_q <= cap_q has coarse sum <= d / unit + 4 <= L + 4: EVERY true top-k node
// passes the test "sum < L + 5".  The survivors (a few hundred per query) are re-scored exactly
// (float tables, double sum, reference arithmetic) by rescore8_kernel, which keeps the k best
// by (distance, position).  A candidate buffer that overflows flags its query for the exact
// fallback.  No result ever depends on the coarse values.
//
// The kernel is scan2's strand design (one 16-byte record per node, four nodes per warp step,
// eight 128-bit reads per node) with the candidate machinery reduced to an append: the bound is
// the constant 36 for every query, so there are no epochs, barriers or compactions.
#include "kernels.cuh"

#include <cfloat>

namespace dpq {

__global__ void cap_from_keys_kernel(const uint64_t* __restrict__ keys, int topk, int Q, float* __restrict__ cap) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) cap[q] = __uint_as_float((uint32_t)(keys[(size_t)q * topk + topk - 1] >> 32));
}
void launch_cap_from_keys(const uint64_t* d_keys, int topk, int Q, float* d_cap, cudaStream_t st) {
    cap_from_keys_kernel<<<(Q + 255) / 256, 256, 0, st>>>(d_keys, topk, Q, d_cap);
}

// One block per query: exact distances of R evenly strided nodes (the query's float table in
// shared memory), then the k-th smallest by bisection on the float bit patterns (distances are
// non-negative, so the integer order of the bits is the float order).
constexpr int PS_T = 128;
__global__ void __launch_bounds__(PS_T) presample_kernel(const float* __restrict__ lutf, const uint8_t* __restrict__ codes,
                                                         int64_t n_local, int M, int K, int topk, int R,
                                                         float* __restrict__ cap) {
    extern __shared__ float s_lut[];  // M*K
    __shared__ int s_cnt[PS_T / 32];
    const int q = blockIdx.x;
    const int MK = M * K;
    for (int i = threadIdx.x; i < MK; i += PS_T) s_lut[i] = lutf[(size_t)q * MK + i];
    __syncthreads();
    constexpr int PER = 16;  // R <= PS_T * PER
    uint32_t v[PER];
    const int64_t stride = n_local / R > 0 ? n_local / R : 1;
#pragma unroll
    for (int t = 0; t < PER; ++t) {
        const int i = t * PS_T + threadIdx.x;
        v[t] = 0x7F800000u;  // +inf: never counted
        if (i < R && (int64_t)i * stride < n_local) {
            const uint8_t* c = codes + (size_t)((int64_t)i * stride) * M;
            double d = 0.0;
            for (int m = 0; m < M; ++m) d += (double)s_lut[m * K + c[m]];
            v[t] = __float_as_uint((float)d);
        }
    }
    // smallest x with count(v <= x) >= topk
    uint32_t lo = 0, hi = 0x7F7FFFFFu;  // FLT_MAX
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int c = 0;
#pragma unroll
        for (int t = 0; t < PER; ++t) c += v[t] <= mid;
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < PS_T / 32; ++w) tot += s_cnt[w];
        __syncthreads();
        if (tot >= topk) hi = mid;
        else lo = mid + 1;
    }
    if (threadIdx.x == 0) cap[q] = __uint_as_float(lo);  // FLT_MAX when the sample holds fewer than k nodes
}
void launch_presample(const float* d_lutf, const uint8_t* d_codes, int64_t n_local, int M, int K, int Q, int topk,
                      int R, float* d_cap, cudaStream_t st) {
    if (R > PS_T * 16) R = PS_T * 16;
    const size_t sm = (size_t)M * K * sizeof(float);
    presample_kernel<<<Q, PS_T, sm, st>>>(d_lutf, d_codes, n_local, M, K, topk, R, d_cap);
}

template <int QB, int ROWS, int SAT>
__global__ void __launch_bounds__(256) pack8_kernel(const float* __restrict__ lutf, const float* __restrict__ capv,
                                                    int MK, int Q, int levels, uint8_t* __restrict__ qlut8,
                                                    uint32_t* __restrict__ ovf) {
    __shared__ uint8_t tile[64 * QB];
    __shared__ double s_inv[QB];
    const int grp = blockIdx.x, row0 = blockIdx.y * 64;
    if (threadIdx.x < QB) {
        const int q = grp * QB + threadIdx.x;
        double inv = 0.0;
        if (q < Q) {
            const float cap = capv[q];
            // cap == FLT_MAX: the sample held fewer than k nodes; every entry quantises to 0 and
            // every node becomes a candidate (correct, merely slow: tiny trees only)
            inv = cap > 0.0f ? (double)levels / (double)cap : 0.0;
            if (cap == 0.0f) inv = 1e30;  // k exact matches: only zero entries stay below the bound
        }
        s_inv[threadIdx.x] = inv;
        if (blockIdx.y == 0) ovf[grp * QB + threadIdx.x] = 0u;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * QB; i += blockDim.x) {
        const int ql = i >> 6, r = i & 63;
        const int q = grp * QB + ql, row = row0 + r;
        uint8_t v = 0;
        if (q < Q && row < MK) {
            const double x = (double)lutf[(size_t)q * MK + row] * s_inv[ql];
            v = x >= (double)SAT ? (uint8_t)SAT : (uint8_t)__double2int_rn(x);
        }
        tile[r * QB + ql] = v;
    }
    __syncthreads();
    uint32_t* dst = reinterpret_cast<uint32_t*>(qlut8 + ((size_t)grp * ROWS + row0) * QB);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tile);
    for (int i = threadIdx.x; i < 64 * QB / 4; i += blockDim.x) dst[i] = src[i];
}

void launch_pack8(const float* d_lutf, const float* d_cap, int MK, int Q, int levels, uint8_t* d_qlut8,
                  uint32_t* d_ovf, int n_groups, int nf, cudaStream_t st) {
    if (nf == 8)
        pack8_kernel<112, 2048, 31><<<dim3((unsigned)n_groups, 2048 / 64), 256, 0, st>>>(d_lutf, d_cap, MK, Q, levels, d_qlut8, d_ovf);
    else
        pack8_kernel<48, 4096, 15><<<dim3((unsigned)n_groups, 4096 / 64), 256, 0, st>>>(d_lutf, d_cap, MK, Q, levels, d_qlut8, d_ovf);
}

// NF fields per record, LPG active 16-byte lanes per strand, SW lanes per strand (as scan2_kernel).
template <int NF, int LPG, int SW>
__global__ void __launch_bounds__(768, 1) scan8_kernel(const Scan8Args a) {
    constexpr int QB = LPG * 16;                 // queries per CTA
    constexpr int ROWS = NF == 8 ? 2048 : 4096;
    constexpr int LUT_BYTES = ROWS * QB;
    constexpr int SPW = 32 / SW;                 // strands per warp = chunks in flight per warp
    constexpr int RW = NF / 8;                   // uint4 words per record
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem + LUT_BYTES);  // [128] candidates per query
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_cnt + 128);

    const int item = blockIdx.x;
    const int slice = item / a.n_groups, grp = item % a.n_groups;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strand = lane / SW, j = lane % SW;
    const int jj = j < LPG ? j : LPG - 1;
    const int n_bt = ((a.n_chunks + SPW - 1) / SPW + a.bt_stride - 1) / a.bt_stride;  // batches this launch walks
    const int b_lo = (int)((int64_t)n_bt * slice / a.n_slices);
    const int b_hi = (int)((int64_t)n_bt * (slice + 1) / a.n_slices);

    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // TMA bulk copy of the group's coarse table
        mbar_expect_tx(s_bar, (uint32_t)LUT_BYTES);
        const unsigned char* src = a.qlut8 + (size_t)grp * LUT_BYTES;
        for (uint32_t o = 0; o < (uint32_t)LUT_BYTES; o += 32768u) bulk_g2s(smem + o, src + o, 32768u, s_bar);
    }
    if (threadIdx.x < 128) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    mbar_wait(s_bar, 0);

    // my 16 queries: ql = jj*16 + 4k + b (word k, byte b).  Per byte: hit iff sum < thresh (<= 128),
    // tested as bit 7 of (0x80 + thresh - 1 - low7(sum)) with bit 7 of the sum clear; a dead
    // byte (idle lane / query beyond Q) gets the constant 0x7F, which never sets bit 7.
    uint32_t lut_base = smem_u32(smem) + (uint32_t)jj * 16u;
    asm volatile("" : "+r"(lut_base));
    const int n_live = j < LPG ? min(16, max(0, a.Q - (grp * QB + jj * 16))) : 0;
    uint32_t cmpc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        cmpc[k] = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) cmpc[k] |= (4 * k + b < n_live ? (0x80u + (uint32_t)a.thresh - 1u) : 0x7Fu) << (8 * b);
    }
    uint32_t* my_cand = a.cand + (size_t)item * QB * a.bcap;
    const int n_rounds = (b_hi - b_lo + a.n_warps - 1) / a.n_warps;
    const int C = a.chunk_nodes;
    uint32_t parp[4] = {1u, 1u, 1u, 1u};

    for (int round = 0; round < n_rounds; ++round) {
        const int bt = b_lo + round * a.n_warps + warp;
        const int c = bt * a.bt_stride * SPW + strand;
        int n_nodes = 0;
        uint32_t pos = 0, rix = 0;
        if (bt < b_hi && c < a.n_chunks) {
            const ChunkDesc2 cd = a.chunks[c];
            n_nodes = (int)cd.n_nodes;
            pos = cd.first_pos;
            rix = cd.rec_begin * RW;
        }
        uint4 rec[RW], nxt[RW];
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            rec[w] = make_uint4(0, 0, 0, 0);
            if (n_nodes > 0) rec[w] = __ldg(a.recs + rix + w);
        }
#pragma unroll 1
        for (int it = 0; it < C; ++it) {
            rix += (uint32_t)RW;
            if ((it & 7) == 0 && it + 25 < n_nodes) {
                prefetch_l2(a.recs + rix + 24 * RW);
                prefetch_l1(a.recs + rix + 8 * RW);
            }
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                nxt[w] = make_uint4(0, 0, 0, 0);
                if (it + 1 < n_nodes) nxt[w] = __ldg(a.recs + rix + w);
            }
            const uint32_t dm = (rec[0].x & V2_ABS) ? 0u : 0xFFFFFFFFu;
            uint32_t d[4];
            if (NF == 8) {
                const uint4 P0 = lds128(fld(lut_base, rec[0].x & 0x3FFFu));
                const uint4 P1 = lds128(fld(lut_base, rec[0].x >> 16));
                const uint4 P2 = lds128(fld(lut_base, rec[0].y & 0xFFFFu));
                const uint4 P3 = lds128(fld(lut_base, rec[0].y >> 16));
                const uint4 M0 = lds128(fld(lut_base, rec[0].z & 0xFFFFu));
                const uint4 M1 = lds128(fld(lut_base, rec[0].z >> 16));
                const uint4 M2 = lds128(fld(lut_base, rec[0].w & 0xFFFFu));
                const uint4 M3 = lds128(fld(lut_base, rec[0].w >> 16));
                d[0] = (P0.x + P1.x + P2.x) + (P3.x + (parp[0] & dm)) + ((M0.x + M1.x + M2.x + M3.x) ^ dm);
                d[1] = (P0.y + P1.y + P2.y) + (P3.y + (parp[1] & dm)) + ((M0.y + M1.y + M2.y + M3.y) ^ dm);
                d[2] = (P0.z + P1.z + P2.z) + (P3.z + (parp[2] & dm)) + ((M0.z + M1.z + M2.z + M3.z) ^ dm);
                d[3] = (P0.w + P1.w + P2.w) + (P3.w + (parp[3] & dm)) + ((M0.w + M1.w + M2.w + M3.w) ^ dm);
            } else {
                // sixteen reads: plus fields = first record word, minus fields = second
                const uint4 rp = rec[0], rm = rec[RW - 1];
                uint32_t sp[4], sm[4];
                {
                    const uint4 A0 = lds128(fld(lut_base, rp.x & 0x3FFFu)), A1 = lds128(fld(lut_base, rp.x >> 16));
                    const uint4 A2 = lds128(fld(lut_base, rp.y & 0xFFFFu)), A3 = lds128(fld(lut_base, rp.y >> 16));
                    const uint4 A4 = lds128(fld(lut_base, rp.z & 0xFFFFu)), A5 = lds128(fld(lut_base, rp.z >> 16));
                    const uint4 A6 = lds128(fld(lut_base, rp.w & 0xFFFFu)), A7 = lds128(fld(lut_base, rp.w >> 16));
                    sp[0] = (A0.x + A1.x + A2.x) + (A3.x + A4.x + A5.x) + (A6.x + A7.x);
                    sp[1] = (A0.y + A1.y + A2.y) + (A3.y + A4.y + A5.y) + (A6.y + A7.y);
                    sp[2] = (A0.z + A1.z + A2.z) + (A3.z + A4.z + A5.z) + (A6.z + A7.z);
                    sp[3] = (A0.w + A1.w + A2.w) + (A3.w + A4.w + A5.w) + (A6.w + A7.w);
                }
                {
                    const uint4 B0 = lds128(fld(lut_base, rm.x & 0xFFFFu)), B1 = lds128(fld(lut_base, rm.x >> 16));
                    const uint4 B2 = lds128(fld(lut_base, rm.y & 0xFFFFu)), B3 = lds128(fld(lut_base, rm.y >> 16));
                    const uint4 B4 = lds128(fld(lut_base, rm.z & 0xFFFFu)), B5 = lds128(fld(lut_base, rm.z >> 16));
                    const uint4 B6 = lds128(fld(lut_base, rm.w & 0xFFFFu)), B7 = lds128(fld(lut_base, rm.w >> 16));
                    sm[0] = (B0.x + B1.x + B2.x) + (B3.x + B4.x + B5.x) + (B6.x + B7.x);
                    sm[1] = (B0.y + B1.y + B2.y) + (B3.y + B4.y + B5.y) + (B6.y + B7.y);
                    sm[2] = (B0.z + B1.z + B2.z) + (B3.z + B4.z + B5.z) + (B6.z + B7.z);
                    sm[3] = (B0.w + B1.w + B2.w) + (B3.w + B4.w + B5.w) + (B6.w + B7.w);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = sp[k] + (parp[k] & dm) + (sm[k] ^ dm);
            }
            const uint32_t am = it < n_nodes ? 0x80808080u : 0u;
            const uint32_t h0 = (cmpc[0] - (d[0] & 0x7F7F7F7Fu)) & ~d[0];
            const uint32_t h1 = (cmpc[1] - (d[1] & 0x7F7F7F7Fu)) & ~d[1];
            const uint32_t h2 = (cmpc[2] - (d[2] & 0x7F7F7F7Fu)) & ~d[2];
            const uint32_t h3 = (cmpc[3] - (d[3] & 0x7F7F7F7Fu)) & ~d[3];
            const uint32_t hit = (h0 | h1 | h2 | h3) & am;
            if (__any_sync(0xffffffffu, hit != 0u)) {
                // rare path: one bit per byte below the bound -> append the position
                const uint32_t hh[4] = {h0 & am, h1 & am, h2 & am, h3 & am};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t m = hh[k];
                    while (m) {
                        const int b = (__ffs(m) - 1) >> 3;
                        m &= m - 1;
                        const int ql = jj * 16 + 4 * k + b;
                        const uint32_t slot = atomicAdd(&s_cnt[ql], 1u);
                        if (slot < (uint32_t)a.bcap) __stcg(my_cand + (size_t)ql * a.bcap + slot, pos);
                        else a.ovf[(size_t)grp * QB + ql] = 1u;
                    }
                }
            }
            if (rec[0].x & V2_CHILD) {
                parp[0] = d[0] + 1u;
                parp[1] = d[1] + 1u;
                parp[2] = d[2] + 1u;
                parp[3] = d[3] + 1u;
            }
#pragma unroll
            for (int w = 0; w < RW; ++w) rec[w] = nxt[w];
            ++pos;
        }
    }
    __syncthreads();
    for (int ql = threadIdx.x; ql < QB; ql += blockDim.x)
        a.cand_cnt[(size_t)item * QB + ql] = min(s_cnt[ql], (uint32_t)a.bcap);
}

cudaError_t launch_scan8(const Scan8Args& a, cudaStream_t st) {
    const C8Shape sh = c8_shape(a.nf);
    const size_t smem = (size_t)sh.lut_bytes() + 128 * 4 + 16;
    if (a.nf == 8) {
        cudaError_t e = cudaFuncSetAttribute(scan8_kernel<8, 7, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        scan8_kernel<8, 7, 8><<<(unsigned)(a.n_groups * a.n_slices), (unsigned)(a.n_warps * 32), smem, st>>>(a);
    } else {
        cudaError_t e = cudaFuncSetAttribute(scan8_kernel<16, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        scan8_kernel<16, 3, 4><<<(unsigned)(a.n_groups * a.n_slices), (unsigned)(a.n_warps * 32), smem, st>>>(a);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------ exact re-score --
// One warp per query: every coarse survivor is scored exactly (float tables, double sum of the
// node's M entries = what the reference's double accumulation rounds to) and the k best keys
// (distance bits << 32 | position) are kept in a small shared-memory buffer that is reduced by
// rank counting whenever it fills.
constexpr int R8_WARPS = 4;
constexpr int R8_BUF = 192;  // >= topk + 32 (coarse search serves topk <= 128); rank counting is O(n^2 / 32)

__device__ __forceinline__ int r8_compact(uint64_t* buf, int n, int k, int lane) {
    // keep the min(n, k) smallest of buf[0..n) sorted ascending (keys unique); n <= R8_BUF
    uint64_t mine[R8_BUF / 32];
    int rank[R8_BUF / 32];
    const int per = (n + 31) >> 5;
#pragma unroll
    for (int t = 0; t < R8_BUF / 32; ++t) {
        mine[t] = ~0ull;
        rank[t] = 0;
        if (t < per && t * 32 + lane < n) mine[t] = buf[t * 32 + lane];
    }
    __syncwarp();
    for (int i = 0; i < n; ++i) {
        const uint64_t o = buf[i];  // broadcast read
#pragma unroll
        for (int t = 0; t < R8_BUF / 32; ++t) rank[t] += o < mine[t];
    }
    __syncwarp();
    const int keep = n < k ? n : k;
#pragma unroll
    for (int t = 0; t < R8_BUF / 32; ++t)
        if (t < per && t * 32 + lane < n && rank[t] < keep) buf[rank[t]] = mine[t];
    __syncwarp();
    return keep;
}

constexpr int R8_MAXSL = 128;  // slices per group the flattened candidate index can address

__global__ void __launch_bounds__(R8_WARPS * 32) rescore8_kernel(const Rescore8Args a) {
    __shared__ uint64_t s_buf[R8_WARPS][R8_BUF];
    __shared__ uint32_t s_off[R8_WARPS][R8_MAXSL + 1];  // exclusive prefix of the per-slice counts
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * R8_WARPS + w;
    if (q >= a.Q) return;
    const int grp = q / a.qb, ql = q % a.qb;
    const float* lut = a.lutf + (size_t)q * a.M * a.K;
    uint64_t* buf = s_buf[w];
    uint32_t* off = s_off[w];
    // all slices' counts at once (one load per lane), then one flat candidate index space so
    // that every step scores 32 candidates no matter how they are spread over the slices
    uint32_t run = 0;
    for (int s0 = 0; s0 < a.n_slices; s0 += 32) {
        const int s = s0 + lane;
        uint32_t c = 0;
        if (s < a.n_slices) c = a.cand_cnt[((size_t)s * a.n_groups + grp) * a.qb + ql];
        uint32_t incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (s < a.n_slices) off[s] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) off[a.n_slices] = run;
    __syncwarp();
    const int total = (int)run;
    int n = 0;
    uint64_t bound = ~0ull;  // k-th best key so far (exclusive)
    const int k = a.topk;
    for (int i0 = 0; i0 < total; i0 += 32) {
        uint64_t key = ~0ull;
        const int i = i0 + lane;
        if (i < total) {
            int lo = 0, hi = a.n_slices;  // slice s with off[s] <= i < off[s+1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (off[mid] <= (uint32_t)i) lo = mid;
                else hi = mid;
            }
            const size_t item = (size_t)lo * a.n_groups + grp;
            const uint32_t pos = __ldcg(a.cand + (item * a.qb + ql) * (size_t)a.bcap + ((uint32_t)i - off[lo]));
            const uint8_t* code = a.codes + (size_t)((int64_t)pos - a.base_pos) * a.M;
            double d = 0.0;
            for (int m = 0; m < a.M; ++m) d += (double)lut[m * a.K + code[m]];
            key = ((uint64_t)__float_as_uint((float)d) << 32) | pos;
        }
        const bool take = key < bound;
        const uint32_t mk = __ballot_sync(0xffffffffu, take);
        if (mk) {
            if (n + 32 > R8_BUF) {  // make room: reduce to the k best, tighten the bound
                n = r8_compact(buf, n, k, lane);
                if (n == k) bound = buf[k - 1];
                __syncwarp();
            }
            const bool still = take && key < bound;
            const uint32_t mk2 = __ballot_sync(0xffffffffu, still);
            if (still) buf[n + __popc(mk2 & ((1u << lane) - 1u))] = key;
            n += __popc(mk2);
            __syncwarp();
        }
    }
    n = r8_compact(buf, n, k, lane);
    if (a.out_key)
        for (int i = lane; i < k; i += 32)
            a.out_key[(size_t)q * k + i] = i < n ? buf[i] : (((uint64_t)__float_as_uint(FLT_MAX) << 32) | 0xFFFFFFFFull);
    // the k best found are real nodes: their k-th distance bounds the true k-th from above even
    // when candidates were dropped
    const float found = n >= k ? __uint_as_float((uint32_t)(buf[k - 1] >> 32)) : FLT_MAX;
    const float known = a.cap_in ? a.cap_in[q] : FLT_MAX;
    if (lane == 0 && a.cap_out) a.cap_out[q] = fminf(found, known);
    if (lane == 0 && a.flagged) {
        // a dropped candidate (buffer overflow) may hide a true top-k node: exact fallback, bounded
        // by the best k found so far
        a.bound[q] = fminf(found, known);
        if (a.ovf[(size_t)grp * a.qb + ql]) {
            const uint32_t slot = atomicAdd(a.n_flagged, 1u);
            if (slot < (uint32_t)a.max_flagged) a.flagged[slot] = (uint32_t)q;
        }
    }
}

void launch_rescore8(const Rescore8Args& a, cudaStream_t st) {
    rescore8_kernel<<<(a.Q + R8_WARPS - 1) / R8_WARPS, R8_WARPS * 32, 0, st>>>(a);
}

}  // namespace dpq
