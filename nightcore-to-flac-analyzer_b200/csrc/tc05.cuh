// Thin inline-PTX layer for the Blackwell (sm_100a) tensor-core path: mbarriers, 1-D bulk async copies,
// TMEM allocation, tcgen05.mma (kind::tf32, A from TMEM, B from a swizzle-128B K-major shared-memory tile),
// tcgen05.st / tcgen05.ld (32 lanes × 32 bit), tcgen05.commit.  Descriptor bit layouts follow the PTX ISA
// (shared-memory matrix descriptor, instruction descriptor) as also encoded in CUTLASS 4.x
// cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ncfa {
namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a fully converged warp (the warp-uniform way to issue single-thread instructions: addresses and
// descriptors stay in uniform registers instead of being broadcast out of a divergent lane)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded spin: a pipeline bug must become a reported launch failure, never a hung GPU.  `backoff_ns` > 0 puts the
// thread to sleep between polls: a spinning warp competes for issue slots with the warps doing the work (measured on the
// CQT kernel: 60 % of all issued instructions were polls).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, unsigned backoff_ns = 0) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (backoff_ns) __nanosleep(backoff_ns);
    }
    __trap();
}

// Whole-warp wait with ONE polling lane: hundreds of threads spinning on the same mbarrier word serialise in shared
// memory (measured: the per-tile handshake of 512 polling threads cost more than the MMAs it guarded).
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity, unsigned backoff_ns = 0) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity, backoff_ns);
    __syncwarp();
}

// ---- 1-D bulk async copy global → shared, completion on an mbarrier (bytes multiple of 16, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {  // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// this thread's lane ← 16 consecutive 32-bit columns starting at taddr (lane field = 32·(warp % 4))
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_vals(uint32_t taddr, const uint32_t (&v)[8]) { tmem_st8(taddr, v); }
__device__ __forceinline__ void tmem_st_vals(uint32_t taddr, const uint32_t (&v)[16]) { tmem_st16(taddr, v); }
__device__ __forceinline__ void tmem_st_vals(uint32_t taddr, const uint32_t (&v)[32]) {
    uint32_t a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = v[i];
        b[i] = v[16 + i];
    }
    tmem_st16(taddr, a);
    tmem_st16(taddr + 16, b);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// ---- descriptors
// K-major operand tile in shared memory, rows of 128 B (32 tf32), 128-byte swizzle, 8-row atoms of 1024 B.
// bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major), [32,46) stride byte
// offset >> 4 (distance between 8-row atoms = 1024 B), [46,48) version = 1 (sm_100), [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc_k128(const void *tile) {
    const uint64_t a = (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);
    return a | (0ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32 (bits [4,6) = 1), A = B = TF32 (bits [7,10) = [10,13) = 2), both K-major,
// N >> 3 at bits [17,23), M >> 4 at bits [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] · B[smem]^T, one thread issues
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// round-to-nearest fp32 → tf32 (low 13 mantissa bits zero)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace tc05
}  // namespace ncfa
