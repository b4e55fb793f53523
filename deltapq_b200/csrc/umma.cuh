// tcgen05 (5th-generation tensor core) helpers shared by the ground-truth filter (gt_tc.cu) and the
// tensor-core encoder (encode_tc.cu): kind::f16 MMA with bf16 operands in the canonical K-major
// shared-memory layout with the 128-byte swizzle, fp32 accumulators in TMEM.
#pragma once
#include <cstdint>

#include "kernels.cuh"  // smem_u32, mbar_init

namespace dpq {
namespace umma {

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N columns, M = 128 rows
constexpr uint32_t idesc_bf16_m128(int n_cols) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// K-major operand tile with the 128-byte swizzle: row r of the tile is 128 contiguous bytes (64 bf16)
// at r * 128, its 16-byte piece c stored at position c ^ (r & 7); 8-row groups are 1024 bytes apart
// (stride byte offset), the leading byte offset is unused (1).  One MMA k-step (16 elements) advances
// the start address by 32 bytes inside the swizzle atom; tiles are 1024-byte aligned (descriptor
// fields as in cute/arch/mma_sm100_desc.hpp, layout type 2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// byte offset of 16-byte piece `piece` of row `row` inside such a tile
__device__ __forceinline__ int swz_off(int row, int piece) { return row * 128 + ((piece ^ (row & 7)) << 4); }

template <uint32_t IDESC>
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// bounded spin on an mbarrier phase: false = the barrier never fired (reported, never a hang)
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

// this thread's TMEM lane, 32 consecutive fp32 columns from `taddr`
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace umma
}  // namespace dpq
