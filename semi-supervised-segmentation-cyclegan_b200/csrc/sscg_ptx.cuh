// sscg_ptx.cuh — thin inline-PTX layer for sm_100a (B200): mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (MMA / TMEM alloc / TMEM load / commit) and UMMA descriptor builders.
//
// Everything here is device-side plumbing for the implicit-GEMM kernels in conv_igemm.cu and
// conv_wgrad.cu.  Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables (same encodings CUTLASS's cute::UMMA::SmemDescriptor /
// InstrDescriptor unions expose).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace sscg {

// ----------------------------------------------------------------------------------------------
// Error flag: every bounded spin (mbarrier wait) reports here and traps instead of hanging the GPU.
// ----------------------------------------------------------------------------------------------
__device__ unsigned int g_sscg_dev_error = 0;   // 0 = ok; otherwise (code << 16) | blockIdx.x low bits

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug (wrong expect_tx, missing commit) becomes a trap + error code after
// ~2 s instead of a hung GPU.
#ifndef SSCG_WAIT_TIMEOUT_NS
#define SSCG_WAIT_TIMEOUT_NS 2000000000ull
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t code) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0) {
            if (globaltimer_ns() - t0 > SSCG_WAIT_TIMEOUT_NS) {
                atomicCAS(&g_sscg_dev_error, 0u, (code << 16) | (blockIdx.x & 0xffff) | 0x80000000u);
                __threadfence_system();
                __trap();
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// TMA (bulk tensor copies global -> shared, completion on an mbarrier)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM loads, fences
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t = TMEM lane base+t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// UMMA descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64-bit):
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4   [46,48) version = 1 (sm_100)
//   [49,52) base offset = 0           [61,64) layout: 0 none, 2 = SWIZZLE_128B
// K-major, SWIZZLE_128B, 64 bf16 (=128 B) per row: 8-row atoms of 1024 B, SBO = 1024, LBO unused.
// MN-major, SWIZZLE_128B: 64 MN-elements (128 B) x 8 K-rows atoms; SBO = stride between 8-K-row
// groups (1024 B for dense rows), LBO = stride between 64-wide MN atoms.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                         uint32_t base_offset = 0) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>(base_offset & 7u) << 49;   // phase of the 8-row swizzle pattern at the start address
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= static_cast<uint64_t>(1) << 46;   // version
    d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
    return d;
}
// Programmatic dependent launch (see launch_k in sscg_common.cuh): block until the predecessor grid has completed and
// its memory operations are visible / allow the successor grid to be scheduled.  Both are no-ops for a launch without
// the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// General form.  layout: 0 = no swizzle ("interleave": 8-row x 16-byte core matrices; LBO = distance between core matrices
// along K, SBO = distance between 8-row groups along M/N), 2 / 4 / 6 = SWIZZLE_128B / 64B / 32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout,
                                                   uint32_t base_offset = 0) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>(base_offset & 7u) << 49;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= static_cast<uint64_t>(1) << 46;   // version
    d |= static_cast<uint64_t>(layout & 7u) << 61;
    return d;
}
// Instruction descriptor (32-bit) for kind::f16 with bf16 A/B and fp32 D:
//   [4,6) D fmt: 1 = f32   [7,10) A fmt: 1 = bf16   [10,13) B fmt: 1 = bf16
//   [15] A major (0 = K, 1 = MN)   [16] B major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}


// ----------------------------------------------------------------------------------------------
// Reproducible cross-CTA sums.
//
// (1) Binned fixed-point accumulators (InstanceNorm plane sums, forward and backward): a float partial sum p is split
//     exactly into hi = rint(p * 2^8) and lo = rint((p - hi * 2^-8) * 2^56) and both parts are added to 64-bit
//     INTEGER accumulators with red.global.add.u64.  Integer addition is associative, so the totals — and the float
//     decoded from them — do not depend on the order in which CTAs arrive: bit-reproducible without any fence,
//     counter or extra launch (floating-point atomics are not: the earlier version differed from run to run).
//     Resolution 2^-56 (1.4e-17) absolute per partial; |p| < 2^40 and up to 2^15 partials per accumulator keep both
//     words far from overflow (|lo| <= 2^47 per partial).
// (2) Fixed-order reductions through a workspace (split-K weight gradients, loss sums): contributors store partials to
//     slots of their own and bump an arrival counter with a gpu-scope acq_rel atomic; the slots are then summed in
//     slot order (by the last arriver, or slice-wise by every contributor once all have arrived).
// ----------------------------------------------------------------------------------------------
constexpr int kDetWords = 2;                      // 64-bit words per accumulated value
__device__ __forceinline__ void det_red_add(unsigned long long* acc, float p) {
    p = fminf(fmaxf(p, -1.0995116e12f), 1.0995116e12f);            // +-2^40: keeps the integer conversion defined
    const float hi_f = rintf(p * 256.f);
    const float r = p - hi_f * 0.00390625f;                         // exact: the bits of p below 2^-8 (|r| <= 2^-9)
    const long long hi = __float2ll_rn(hi_f);
    const long long lo = __float2ll_rn(r * 72057594037927936.f);    // 2^56
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(acc), "l"(static_cast<unsigned long long>(hi)) : "memory");
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(acc + 1), "l"(static_cast<unsigned long long>(lo)) : "memory");
}
// hi * 2^-8 + lo * 2^-56 as a float: each word as sign and magnitude, the magnitude from its 32-bit halves with fp32
// FMAs (no fp64, no 64-bit conversions).  A pure function of the integers, hence as reproducible as they are;
// relative error ~2^-23 of the larger word's contribution (fp32 level).
__device__ __forceinline__ float det_word(long long x, float scale_hi, float scale_lo) {
    const bool neg = x < 0;
    const unsigned long long m = neg ? static_cast<unsigned long long>(-x) : static_cast<unsigned long long>(x);
    const float v = fmaf(static_cast<float>(static_cast<unsigned int>(m >> 32)), scale_hi,
                         static_cast<float>(static_cast<unsigned int>(m)) * scale_lo);
    return neg ? -v : v;
}
__device__ __forceinline__ float det_decode(long long hi, long long lo) {
    return det_word(hi, 16777216.f, 0.00390625f) +                                 // 2^24, 2^-8
           det_word(lo, 5.9604644775390625e-8f, 1.3877787807814457e-17f);          // 2^-24, 2^-56
}

__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu(unsigned int* addr, unsigned int v) {
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* addr) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
// Bounded spin until *addr >= target (all co-resident contributors have arrived); a protocol bug traps after ~2 s.
__device__ __forceinline__ void spin_until_ge(const unsigned int* addr, unsigned int target, uint32_t code) {
    if (ld_acquire_gpu(addr) >= target) return;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (ld_acquire_gpu(addr) < target) {
        if ((++spins & 0xff) == 0 && globaltimer_ns() - t0 > SSCG_WAIT_TIMEOUT_NS) {
            atomicCAS(&g_sscg_dev_error, 0u, (code << 16) | (blockIdx.x & 0xffff) | 0x80000000u);
            __threadfence_system();
            __trap();
        }
    }
}
__device__ __forceinline__ float2 ld_cg_f2(const float* p) {
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

}  // namespace sscg
