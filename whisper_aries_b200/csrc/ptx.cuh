// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM), fences.
// Hand-written for this project; encodings follow the PTX ISA 8.7 tcgen05 chapter.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace aries {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Every wait is bounded: a protocol bug must surface as a launch failure ("unspecified launch failure" through the C
// ABI), never as a hung GPU.  A legitimate wait lasts microseconds; the bound is on the order of seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// For roles that are not latency critical (TMA producer): back off between probes so the spin does not eat issue
// slots of the compute warps sharing the scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(64);
        if (++spins > (1u << 24)) __trap();
    }
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// Convergent forms for a producer WARP (every lane executes them with warp-uniform operands; the instruction itself is
// predicated on elect.sync).  Under `if (lane == 0)` ptxas wraps each TMA instruction in an ELECT / R2UR.BROADCAST /
// BRA.U.ANY loop; together with an integer division in the K loop that made the GEMM's producer ~75 instructions per
// K block with a MUFU.RCP chain -- too slow to refill a stage in time when two GELU-busy epilogue warps share its
// scheduler (round-2 timeline: 20 % of fc1's main loop was the MMA warp waiting for operands because of it).
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---------------------------------------------------------------- TMEM alloc
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {   // whole warp
    static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {      // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (64 bit): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout [61,64) (2 = SWIZZLE_128B).
__host__ __device__ constexpr uint64_t umma_smem_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16, bf16 x bf16 -> f32.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)            // D format f32
           | (1u << 7)          // A format bf16
           | (1u << 10)         // B format bf16
           | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Same kind::f16 descriptor with f16 x f16 operands (formats 0) -> f32: the LayerNorm-folded GEMMs read the f16 residual
// stream as their A operand and keep gamma-scaled weights in f16.
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// tcgen05.mma / tcgen05.commit are issued by ONE thread.  Every lane of the issuing warp executes the wrappers below
// (with identical, warp-uniform operands) and the instruction itself is predicated on elect.sync.  Under an ordinary
// divergent `if (lane == 0)` ptxas wraps each tcgen05 instruction in a loop over the active lanes (~13 instructions,
// ~80 issue cycles per MMA), which bounded both the attention kernel (small MMAs) and the GEMM.
//
// Arrive on an mbarrier once every previously issued tcgen05.mma of the elected thread has completed.
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// Four K = 16 steps of one 64-deep contraction in a single statement (one elect.sync for all of them).  The shared-
// memory descriptors are passed as (low word, high word): the high word is a per-kernel constant and the low word
// advances by (byte offset >> 4), i.e. by 2 per 32-byte step; a TMEM A operand advances by 8 columns.  Steps 1..3
// always accumulate.  D[tmem] (+)= A[smem] * B[smem]:
__device__ __forceinline__ void umma_bf16_ss_x4_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                                      uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "add.u32 a, %1, 2;\n\tadd.u32 b, %2, 2;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t"
        "add.u32 a, %1, 4;\n\tadd.u32 b, %2, 4;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t"
        "add.u32 a, %1, 6;\n\tadd.u32 b, %2, 6;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows on lanes 0..127, K-major, two bf16 per 32-bit column) is
// read from tensor memory -- attention feeds P straight from the softmax warps' tcgen05.st.
__device__ __forceinline__ void umma_bf16_ts_x4_elect(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo,
                                                      uint32_t desc_hi, uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t.reg .b32 a, b;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
        "add.u32 a, %1, 8;\n\tadd.u32 b, %2, 2;\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], db, %4, 1;\n\t"
        "add.u32 a, %1, 16;\n\tadd.u32 b, %2, 4;\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], db, %4, 1;\n\t"
        "add.u32 a, %1, 24;\n\tadd.u32 b, %2, 6;\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a], db, %4, 1;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
        : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of one cluster (the two SMs of a TPC) execute ONE tcgen05.mma with M = 256: each holds its own 128 rows of
// A and of the accumulator, and HALF of the B tile, which the instruction reads from both shared memories.  Only the
// rank-0 CTA issues MMAs; both issue TMA loads that signal rank 0's mbarrier, and the commit is multicast to both.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address -> rank 0
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the RANK-0 CTA's barrier at the same shared-memory offset.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_elect(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];\n\t}" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the rank-0 CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot) {   // one warp (same index) in BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// Four K = 16 steps, M = 256 over the CTA pair (see umma_bf16_ss_x4_elect).
__device__ __forceinline__ void umma_bf16_ss_x4_elect_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                                           uint32_t desc_hi, uint32_t idesc, uint32_t accumulate_first) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
        "add.u32 a, %1, 2;\n\tadd.u32 b, %2, 2;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, 1;\n\t"
        "add.u32 a, %1, 4;\n\tadd.u32 b, %2, 4;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, 1;\n\t"
        "add.u32 a, %1, 6;\n\tadd.u32 b, %2, 6;\n\tmov.b64 da, {a, %3};\n\tmov.b64 db, {b, %3};\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
        : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_elect_pair(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mov.b16 m, 3;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base+i), r[j] = column (base+j).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
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
// tcgen05.wait::ld that also names the destination registers of the load it completes, so that the compiler cannot
// schedule a use of r[] above the wait (the load is asynchronous; the registers are only valid after it).
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// registers -> TMEM, same 32x32b.x32 shape as the load
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// 16-column variants (32 lanes x 16 consecutive 32-bit columns)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 256-bit global accesses (sm_100): one full 32-byte sector per thread and instruction.
__device__ __forceinline__ void stg256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                       uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
                 "r"(f), "r"(g), "r"(h)
                 : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t& e,
                                       uint32_t& f, uint32_t& g, uint32_t& h) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h)
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void ldg256_nc(const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t& e,
                                          uint32_t& f, uint32_t& g, uint32_t& h) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h)
                 : "l"(p));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// GELU (exact-erf definition) on two values at once, on the FMA pipe only: x * Phi(x) with Phi(x) = 0.5 + xc Q(xc^2),
// xc = clamp(x, +-4.25), Q a degree-8 minimax-style (Chebyshev) fit of (Phi(x) - 0.5) / x.  |error| <= 9.1e-5 absolute
// over all x (checked in f32 against erf in f64; tests/test_gelu_poly.py) -- two orders below the bf16 resolution of the
// outputs it produces -- with packed f32x2 instructions: ~8 issue slots and no MUFU per element, against ~15 + 2 MUFU
// for gelu_fast().  The fc1 epilogue, the largest GEMM's, is bounded by its instruction count.
__device__ __forceinline__ float2 gelu_poly2(float2 x) {
    const float kX0 = 4.25f;
    const float2 xm = make_float2(fmaxf(x.x, -kX0), fmaxf(x.y, -kX0));     // (also the multiplier: keeps x Phi bounded
    const float2 xc = make_float2(fminf(xm.x, kX0), fminf(xm.y, kX0));     //  below -4.25, where Phi is not exactly 0)
    const float2 u = __fmul2_rn(xc, xc);
    float2 q = make_float2(6.7213117016518e-11f, 6.7213117016518e-11f);
    q = __ffma2_rn(q, u, make_float2(-6.215662651243292e-09f, -6.215662651243292e-09f));
    q = __ffma2_rn(q, u, make_float2(2.5361430289194686e-07f, 2.5361430289194686e-07f));
    q = __ffma2_rn(q, u, make_float2(-6.0973584368184675e-06f, -6.0973584368184675e-06f));
    q = __ffma2_rn(q, u, make_float2(9.791838238015771e-05f, 9.791838238015771e-05f));
    q = __ffma2_rn(q, u, make_float2(-0.001132951583713293f, -0.001132951583713293f));
    q = __ffma2_rn(q, u, make_float2(0.009886087849736214f, 0.009886087849736214f));
    q = __ffma2_rn(q, u, make_float2(-0.06643500179052353f, -0.06643500179052353f));
    q = __ffma2_rn(q, u, make_float2(0.3989364206790924f, 0.3989364206790924f));
    const float2 phi = __ffma2_rn(xc, q, make_float2(0.5f, 0.5f));
    return __fmul2_rn(xm, phi);
}

}  // namespace aries
