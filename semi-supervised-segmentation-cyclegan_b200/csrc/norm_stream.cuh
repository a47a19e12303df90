// norm_stream.cuh — bulk-copy pipelined variants of the three InstanceNorm passes (fast bf16 mode).
//
// The register-batched kernels in norm_kernels.cuh top out at 2-4 TB/s: the bytes a thread can keep in
// flight are bounded by its registers, and the load -> compute -> store phases of a batch do not
// overlap.  Here the in-flight bytes live in shared memory instead: one elected thread streams
// contiguous NHWC segments ("units": part of one image row, <= 16 KB per tensor) into a ring of
// stages with cp.async.bulk (TMA's 1-D bulk copy, completion on an mbarrier), 100-190 KB per SM in
// flight, while all 512 threads consume the previous stage out of shared memory and write their
// results with 16-byte global stores.  One persistent CTA per SM walks a contiguous range of units, so
// per-(sample, channel) partial sums stay in registers and are flushed once per sample change.
//
// Arithmetic, rounding points and the dropout hash are identical to norm_kernels.cuh (the parity-mode
// and odd-shape paths keep using those kernels).  Included by elementwise.cu inside namespace sscg.
#pragma once
#include <type_traits>

#ifndef SSCG_STR_UNIT_KB
#define SSCG_STR_UNIT_KB 32
#endif
#ifndef SSCG_STR_BUDGET_KB
#define SSCG_STR_BUDGET_KB 192
#endif
// consumer threads per CTA (one more warp produces): 512 for the light kernels (<= 96 registers), 448 for the
// first backward half, whose two-pixel batches and per-channel accumulators need ~128 registers
constexpr int kStrThreadsLight = 512;
constexpr int kStrThreadsHeavy = 448;
constexpr int kStrMaxStages = 8;
constexpr int kStrSmemBudget = SSCG_STR_BUDGET_KB * 1024;   // ring bytes per CTA (one CTA per SM)
constexpr int kStrUnitMax = SSCG_STR_UNIT_KB * 1024;        // preferred upper bound of one unit, per tensor

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct StreamGeom {
    int upr;        // units per image row
    int seg_px;     // pixels per unit
    int ups;        // units per sample = H * upr
    int total;      // N * ups
    int ub;         // bytes per unit per tensor
    int nst;        // ring stages
    int ntens;      // tensors loaded per unit
    int CH;         // 8-channel vectors per pixel
    int slot0;      // bytes reserved for the first tensor of a stage (>= ub: the backward kernel also stages halo columns)
    int stage_bytes;
};

// volatile: the same shared address is re-read after the stage has been refilled; ordered against the
// (volatile) mbarrier wait that precedes it
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void cvt8(const uint4& u, float (&v)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[2 * q] = __uint_as_float(w[q] << 16);
        v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// Ring of stages: full[s] completes when the bulk copies of a unit have landed, empty[s] when every
// consumer warp has finished reading it.  Warp kStrWarps is the producer; consumer warps run freely.
struct StreamRing {
    uint32_t base;          // shared-space address of stage 0
    uint32_t bars;          // shared-space address of full[0]; empty[s] follows the full barriers
    int nst, stage_bytes;
    __device__ __forceinline__ uint32_t stage(int k) const { return base + (k % nst) * stage_bytes; }
    __device__ __forceinline__ uint32_t full(int k) const { return bars + (k % nst) * 8; }
    __device__ __forceinline__ uint32_t empty(int k) const { return bars + (kStrMaxStages + k % nst) * 8; }
    __device__ __forceinline__ uint32_t parity(int k) const { return (uint32_t)((k / nst) & 1); }
    // producer side: stage of unit k is free once the consumers released its previous occupant (unit k - nst)
    __device__ __forceinline__ void acquire(int k) const {
        if (k >= nst) mbar_wait(empty(k), (uint32_t)(((k / nst) - 1) & 1), 14);
    }
    // consumer side: all lanes of a warp are done with the stage of unit k
    __device__ __forceinline__ void release(int k) const {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty(k));
    }
};

template <int kStrThreads>
__device__ __forceinline__ StreamRing stream_ring_init(uint8_t* smem_raw, const StreamGeom& g) {
    constexpr int kStrWarps = kStrThreads / 32;
    StreamRing r;
    const uint32_t al = (smem_u32(smem_raw) + 127u) & ~127u;
    r.bars = al;
    r.base = al + 128;
    r.nst = g.nst;
    r.stage_bytes = g.stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < g.nst; ++s) {
            mbar_init(r.bars + s * 8, 1);
            mbar_init(r.bars + (kStrMaxStages + s) * 8, kStrWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    return r;
}

// Padded positions that mirror onto source index s under reflection, as scalars (no local arrays):
// m1 / m2 = -1 when absent.  The position s + pad itself is handled by the caller.
__device__ __forceinline__ void mirror_pos(int s, int n, int pad, int& m1, int& m2) {
    m1 = (s >= 1 && s <= pad) ? pad - s : -1;
    m2 = (s <= n - 2 && s >= n - 1 - pad) ? pad + 2 * (n - 1) - s : -1;
}

struct UnitPos {
    int n, h, w0;
};
__device__ __forceinline__ UnitPos unit_pos(const StreamGeom& g, int u) {
    UnitPos p;
    p.n = u / g.ups;
    const int r = u - p.n * g.ups;
    p.h = r / g.upr;
    p.w0 = (r - p.h * g.upr) * g.seg_px;
    return p;
}

// ---------------------------------------------------------------------------------------------
// forward: y = dropout(act(instance_norm(raw))) (+ residual), written with (reflect) halo
// ---------------------------------------------------------------------------------------------
struct ApplyStreamDev {
    SscgApplyArgs a;
    StreamGeom g;
};

template <int kStrThreads>
__global__ void __launch_bounds__(kStrThreads + 32, 1) in_apply_stream_kernel(const __grid_constant__ ApplyStreamDev p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float2 s_pairs[kMaxStreamC];
    const SscgApplyArgs& a = p.a;
    const StreamGeom& g = p.g;
    const StreamRing ring = stream_ring_init<kStrThreads>(smem_raw, g);
    const int u0 = (int)((long long)blockIdx.x * g.total / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * g.total / gridDim.x);
    const int cnt = u1 - u0;
    if (cnt <= 0) return;
    const bool norm = a.stats != nullptr;
    const bool has_res = a.res.ptr != nullptr;
    if (threadIdx.x >= kStrThreads) {
        // ================================ producer warp ========================================
        if (threadIdx.x == kStrThreads) {
            const __nv_bfloat16* rawp = reinterpret_cast<const __nv_bfloat16*>(a.raw);
            const __nv_bfloat16* resp = reinterpret_cast<const __nv_bfloat16*>(a.res.ptr);
            for (int k = 0; k < cnt; ++k) {
                const UnitPos up = unit_pos(g, u0 + k);
                ring.acquire(k);
                const uint32_t bar = ring.full(k), dst = ring.stage(k);
                mbar_arrive_expect_tx(bar, (uint32_t)(g.ntens * g.ub));
                bulk_load(dst, rawp + (((long long)up.n * a.H + up.h) * a.W + up.w0) * a.C, g.ub, bar);
                if (has_res)
                    bulk_load(dst + g.ub, resp + (long long)up.n * a.res.sN + (long long)up.h * a.res.sH +
                                              (long long)up.w0 * a.res.sW, g.ub, bar);
            }
        }
        return;
    }
    // ==================================== consumers ============================================
    const int chunk = threadIdx.x % g.CH;
    const int c0 = chunk * 8;
    const int prow = threadIdx.x / g.CH;            // first pixel of this thread inside a unit
    const int pstep = kStrThreads / g.CH;
    const int Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad;
    const bool reflect = a.pad > 0 && a.pad_mode == SSCG_PAD_REFLECT;
    const uint64_t seed = (a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                           : a.drop_seed;
    float mean[8], rstd[8];
    int cur_n = -1;
    for (int k = 0; k < cnt; ++k) {
        const UnitPos up = unit_pos(g, u0 + k);
        const int n = up.n, h = up.h, w0 = up.w0;
        if (n != cur_n) {
            cur_n = n;
            if (norm) cta_load_sums<false>(a.stats, n, a.C, 1.f / (float)(a.H * a.W), a.eps, s_pairs, threadIdx.x, kStrThreads,
                                           c0, true, BarNamed1<kStrThreads>(), mean, rstd);
        }
        int hm1 = -1, hm2 = -1;
        if (reflect) mirror_pos(h, a.H, a.pad, hm1, hm2);
        const bool hmirror = (hm1 >= 0) || (hm2 >= 0);
        const long long spix0 = ((long long)n * a.H + h) * a.W;
        mbar_wait(ring.full(k), ring.parity(k), 11);
        const uint32_t st = ring.stage(k);
        // per-unit pointers (64-bit arithmetic once per unit); a pixel then costs one 32-bit multiply
        const uint32_t pxb = (uint32_t)a.C * 2u;
        const uint32_t s_v0 = st + (uint32_t)c0 * 2u;
        uint8_t* d0 = reinterpret_cast<uint8_t*>(a.dst) + ((((long long)n * Hp + (h + a.pad)) * Wp + (w0 + a.pad)) * a.C + c0) * 2;
        const uint32_t vec0 = (uint32_t)((spix0 + w0) * g.CH + chunk);
        // does this unit hold a pixel that is mirrored into the halo?  (uniform per unit)
        const bool border_unit = reflect && (hmirror || w0 <= a.pad || w0 + g.seg_px - 1 >= a.W - 1 - a.pad);
        auto body = [&](auto border_tag) {
            constexpr bool BORDER = decltype(border_tag)::value;
            for (int px0 = prow; px0 < g.seg_px; px0 += 2 * pstep) {
                uint4 rv[2], rr[2];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int px = px0 + b * pstep;
                    if (px < g.seg_px) {
                        const uint32_t so = s_v0 + (uint32_t)px * pxb;
                        rv[b] = lds128(so);
                        if (has_res) rr[b] = lds128(so + g.ub);
                    }
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int px = px0 + b * pstep;
                    if (px >= g.seg_px) continue;
                    float v[8];
                    cvt8(rv[b], v);
                    if (norm) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = (v[q] - mean[q]) * rstd[q];
                    }
                    if (a.act == SSCG_ACT_RELU) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
                    } else if (a.act == SSCG_ACT_LRELU) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * a.slope;
                    }
                    if (seed != 0) {
                        const uint32_t bits = drop_bits(seed, vec0 + (uint32_t)px * (uint32_t)g.CH);
                        // keep -> x2, drop -> x0: bit q moved to the exponent position of 2.0f (0x40000000)
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] *= __uint_as_float((bits << (30 - q)) & 0x40000000u);
                    }
                    if (has_res) {
                        float r8[8];
                        cvt8(rr[b], r8);
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] += r8[q];
                    }
                    const uint4 out = pack8(v);
                    *reinterpret_cast<uint4*>(d0 + (uint32_t)px * pxb) = out;
                    if (BORDER) {
                        // halo copies: only pixels within `pad` of a border have mirror positions
                        const int w = w0 + px;
                        if (hmirror || w <= a.pad || w >= a.W - 1 - a.pad) {
                            int wm1, wm2;
                            mirror_pos(w, a.W, a.pad, wm1, wm2);
                            __nv_bfloat16* dbase = reinterpret_cast<__nv_bfloat16*>(a.dst) + (long long)n * Hp * Wp * a.C + c0;
                            const int hh[3] = {h + a.pad, hm1, hm2}, ww[3] = {w + a.pad, wm1, wm2};
#pragma unroll
                            for (int x = 0; x < 3; ++x)
#pragma unroll
                                for (int y = 0; y < 3; ++y)
                                    if (x + y > 0 && hh[x] >= 0 && ww[y] >= 0)
                                        *reinterpret_cast<uint4*>(dbase + ((long long)hh[x] * Wp + ww[y]) * a.C) = out;
                        }
                    }
                }
            }
        };
        if (border_unit) body(std::true_type{});
        else body(std::false_type{});
        ring.release(k);
    }
}

// ---------------------------------------------------------------------------------------------
// backward, first half: dZ, optional folded total gradient, per-(n, c) sums
// ---------------------------------------------------------------------------------------------
struct BwdStreamDev {
    SscgBwdArgs a;
    void* draw;
    StreamGeom g;
};

// consumers only (named barrier 1 over kStrThreads threads).  s_red: kStrThreads * 16 floats.  Thread
// (row, chunk) parks its 16 partial sums in row `row`; CH * 16 outputs are then summed over the rows.
// (A first version used shared-memory float atomics: kStrThreads / CH-way contended CAS loops, ~6 us per
// flush — as much as the whole streaming part of the kernel.)
template <int kStrThreads>
__device__ __forceinline__ void stream_flush_stats(float* s_red, float (&acc1)[8], float (&acc2)[8], void* bstats,
                                                   int n, int C, int CH, int chunk) {
    const int row = threadIdx.x / CH, rows = kStrThreads / CH, width = CH * 16;
    float4* mine = reinterpret_cast<float4*>(s_red + row * width + chunk * 16);
    mine[0] = make_float4(acc1[0], acc2[0], acc1[1], acc2[1]);
    mine[1] = make_float4(acc1[2], acc2[2], acc1[3], acc2[3]);
    mine[2] = make_float4(acc1[4], acc2[4], acc1[5], acc2[5]);
    mine[3] = make_float4(acc1[6], acc2[6], acc1[7], acc2[7]);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc1[q] = acc2[q] = 0.f;
    named_bar_sync(1, kStrThreads);
    for (int o = threadIdx.x; o < width; o += kStrThreads) {
        float t = 0.f;
        for (int r = 0; r < rows; ++r) t += s_red[r * width + o];
        // o = ch * 16 + 2 * q + k  ->  accumulator of bstats[n][ch * 8 + q][k]: order-independent integer sums (sscg_ptx.cuh)
        det_red_add(reinterpret_cast<unsigned long long*>(bstats) + ((long long)n * C * 2 + o) * kDetWords, t);
    }
    named_bar_sync(1, kStrThreads);
}

// SPEC: 0 = every flag read at run time; 1 / 2 = the two residual-block shapes with their flags folded at
// compile time (1: norm + ReLU, no skip / total-gradient output — conv1 of a block; 2: norm, no activation,
// skip gradient + total-gradient output, no dropout — conv2 of a block).  The generic kernel spends ~70 of its
// ~240 instructions per 8-channel vector on uniform flag tests and their predication.
template <int kStrThreads, int SPEC>
__global__ void __launch_bounds__(kStrThreads + 32, 1) in_bwd_prep_stream_kernel(const __grid_constant__ BwdStreamDev p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(16) float s_red[kStrThreads * 16];
    __shared__ float2 s_pairs[kMaxStreamC];
    const SscgBwdArgs& a = p.a;
    const StreamGeom& g = p.g;
    const StreamRing ring = stream_ring_init<kStrThreads>(smem_raw, g);
    const int u0 = (int)((long long)blockIdx.x * g.total / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * g.total / gridDim.x);
    const int cnt = u1 - u0;
    if (cnt <= 0) return;
    const bool norm = SPEC != 0 ? true : (a.stats != nullptr);
    const int act = SPEC == 1 ? SSCG_ACT_RELU : (SPEC == 2 ? SSCG_ACT_NONE : a.act);
    const bool need_raw = norm || act != SSCG_ACT_NONE;
    const bool has_skip = SPEC == 1 ? false : (SPEC == 2 ? true : (a.skip.ptr != nullptr));
    const bool has_gout = SPEC == 1 ? false : (SPEC == 2 ? true : (a.g_out != nullptr));
    const bool fold = (a.pad_mode == SSCG_PAD_REFLECT) && a.pad > 0;
    const int off_skip = g.slot0, off_raw = g.slot0 + (has_skip ? 1 : 0) * g.ub;
    // The gradient row is staged WITH the halo columns next to it (first / last unit of a row), so the
    // mirrored columns of a reflect halo fold out of shared memory; only the 2 * pad rows next to the top /
    // bottom border still fetch their mirrored rows from global memory.
    const int last_part_w0 = (g.upr - 1) * g.seg_px;
    if (threadIdx.x >= kStrThreads) {
        if (threadIdx.x == kStrThreads) {
            const __nv_bfloat16* rawp = reinterpret_cast<const __nv_bfloat16*>(a.raw);
            const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(a.dyp.ptr);
            const __nv_bfloat16* skp = reinterpret_cast<const __nv_bfloat16*>(a.skip.ptr);
            for (int k = 0; k < cnt; ++k) {
                const UnitPos up = unit_pos(g, u0 + k);
                ring.acquire(k);
                const uint32_t bar = ring.full(k), dst = ring.stage(k);
                const int lpad = (fold && up.w0 == 0) ? a.pad : 0;
                const int rpad = (fold && up.w0 == last_part_w0) ? a.pad : 0;
                const uint32_t dy_bytes = (uint32_t)((g.seg_px + lpad + rpad) * a.C * 2);
                mbar_arrive_expect_tx(bar, dy_bytes + (uint32_t)((g.ntens - 1) * g.ub));
                bulk_load(dst, dyp + (long long)up.n * a.dyp.sN + (long long)(up.h + a.pad) * a.dyp.sH +
                                   (long long)(up.w0 + a.pad - lpad) * a.dyp.sW, dy_bytes, bar);
                if (has_skip)
                    bulk_load(dst + off_skip, skp + (long long)up.n * a.skip.sN + (long long)up.h * a.skip.sH +
                                                  (long long)up.w0 * a.skip.sW, g.ub, bar);
                if (need_raw)
                    bulk_load(dst + off_raw, rawp + (((long long)up.n * a.H + up.h) * a.W + up.w0) * a.C, g.ub, bar);
            }
        }
        return;
    }
    const int chunk = threadIdx.x % g.CH;
    const int c0 = chunk * 8;
    const int prow = threadIdx.x / g.CH;
    const int pstep = kStrThreads / g.CH;
    const uint64_t seed = SPEC == 2 ? 0ull
                          : ((a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                              : a.drop_seed);
    float mean[8], rstd[8], acc1[8], acc2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc1[q] = acc2[q] = 0.f;
    int cur_n = -1;
    for (int k = 0; k < cnt; ++k) {
        const UnitPos up = unit_pos(g, u0 + k);
        const int n = up.n, h = up.h, w0 = up.w0;
        if (n != cur_n) {
            if (cur_n >= 0 && a.bstats != nullptr) stream_flush_stats<kStrThreads>(s_red, acc1, acc2, a.bstats, cur_n, a.C, g.CH, chunk);
            cur_n = n;
            if (norm) cta_load_sums<false>(a.stats, n, a.C, 1.f / (float)(a.H * a.W), a.eps, s_pairs, threadIdx.x, kStrThreads,
                                           c0, true, BarNamed1<kStrThreads>(), mean, rstd);
        }
        int hm1 = -1, hm2 = -1;
        if (fold) mirror_pos(h, a.H, a.pad, hm1, hm2);
        const bool hborder = (hm1 >= 0) || (hm2 >= 0);
        const long long spix0 = ((long long)n * a.H + h) * a.W;
        const int lpad = (fold && w0 == 0) ? a.pad : 0;
        const int col0 = w0 + a.pad - lpad;            // padded column held by staged pixel 0 of the gradient row
        mbar_wait(ring.full(k), ring.parity(k), 12);
        const uint32_t st = ring.stage(k);
        // Everything that depends on the unit only is resolved here (64-bit pointer arithmetic once per unit); a pixel
        // of the unit then costs one 32-bit multiply for its byte offset.  (The first version recomputed 64-bit element
        // offsets per pixel: integer instructions were 60 % of the kernel, which is issue-bound.)
        const uint32_t pxb = (uint32_t)a.C * 2u;                                   // bytes per pixel
        const uint32_t s_g0 = st + (uint32_t)lpad * pxb + (uint32_t)c0 * 2u;      // gradient of pixel w0
        const uint32_t s_k0 = st + (uint32_t)off_skip + (uint32_t)c0 * 2u;
        const uint32_t s_z0 = st + (uint32_t)off_raw + (uint32_t)c0 * 2u;
        uint8_t* gout0 = reinterpret_cast<uint8_t*>(a.g_out) + ((spix0 + w0) * a.C + c0) * 2;
        uint8_t* dz0 = reinterpret_cast<uint8_t*>(a.dz) +
                       (a.dz_pad > 0
                            ? ((((long long)n * (a.H + 2 * a.dz_pad) + h + a.dz_pad) * (a.W + 2 * a.dz_pad) + w0 + a.dz_pad) * a.C + c0)
                            : ((spix0 + w0) * a.C + c0)) * 2;
        const uint32_t vec0 = (uint32_t)((spix0 + w0) * g.CH + chunk);              // dropout hash index of pixel w0 (mod 2^32)
        // rows next to the top / bottom border: the gradient of the mirrored halo row (same column) is fetched
        // from global memory up front, together with the shared-memory loads of the batch
        const int xf = hm1 >= 0 ? 1 : 2;
        const uint8_t* mrow0 = reinterpret_cast<const uint8_t*>(
            reinterpret_cast<const __nv_bfloat16*>(a.dyp.ptr) + (long long)n * a.dyp.sN +
            (long long)(hm1 >= 0 ? hm1 : (hm2 >= 0 ? hm2 : 0)) * a.dyp.sH + (long long)(w0 + a.pad) * a.dyp.sW + c0);
        // does this unit hold a pixel whose gradient receives mirrored halo contributions?  (uniform per unit)
        const bool border_unit = hborder || (fold && (w0 <= a.pad || w0 + g.seg_px - 1 >= a.W - 1 - a.pad));
        auto body = [&](auto border_tag) {
            constexpr bool BORDER = decltype(border_tag)::value;
            for (int px0 = prow; px0 < g.seg_px; px0 += 2 * pstep) {
                uint4 rg[2], rs[2], rz[2], rm[2];
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int px = px0 + b * pstep;
                    if (px < g.seg_px) {
                        const uint32_t so = (uint32_t)px * pxb;
                        if (BORDER && hborder) rm[b] = *reinterpret_cast<const uint4*>(mrow0 + so);
                        rg[b] = lds128(s_g0 + so);
                        if (has_skip) rs[b] = lds128(s_k0 + so);
                        if (need_raw) rz[b] = lds128(s_z0 + so);
                    }
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int px = px0 + b * pstep;
                    if (px >= g.seg_px) continue;
                    const uint32_t so = (uint32_t)px * pxb;
                    float gv[8], z[8];
                    cvt8(rg[b], gv);
                    if (BORDER) {
                        const int w = w0 + px;
                        if (hborder || (w <= a.pad || w >= a.W - 1 - a.pad)) {
                            int wm1, wm2;
                            mirror_pos(w, a.W, a.pad, wm1, wm2);
                            const int hh[3] = {h + a.pad, hm1, hm2}, ww[3] = {w + a.pad, wm1, wm2};
                            // same accumulation order as fold_positions() in norm_kernels.cuh: (h, w) = base, m1, m2 nested
#pragma unroll
                            for (int x = 0; x < 3; ++x)
#pragma unroll
                                for (int y = 0; y < 3; ++y) {
                                    if (x + y == 0 || hh[x] < 0 || ww[y] < 0) continue;
                                    float t[8];
                                    if (x == 0)       // same row: the mirrored column was staged with the row
                                        cvt8(lds128(st + ((ww[y] - col0) * a.C + c0) * 2), t);
                                    else if (x == xf && y == 0)
                                        cvt8(rm[b], t);
                                    else
                                        load8(a.dyp.ptr, false,
                                              (long long)n * a.dyp.sN + (long long)hh[x] * a.dyp.sH + (long long)ww[y] * a.dyp.sW + c0, t);
#pragma unroll
                                    for (int q = 0; q < 8; ++q) gv[q] += t[q];
                                }
                        }
                    }
                    if (has_skip) {
                        float t[8];
                        cvt8(rs[b], t);
#pragma unroll
                        for (int q = 0; q < 8; ++q) gv[q] += t[q];
                    }
                    if (has_gout) *reinterpret_cast<uint4*>(gout0 + so) = pack8(gv);
                    if (seed != 0) {
                        const uint32_t bits = drop_bits(seed, vec0 + (uint32_t)px * (uint32_t)g.CH);
                        // keep -> x2, drop -> x0: bit q moved to the exponent position of 2.0f (0x40000000)
#pragma unroll
                        for (int q = 0; q < 8; ++q) gv[q] *= __uint_as_float((bits << (30 - q)) & 0x40000000u);
                    }
                    if (need_raw) {
                        cvt8(rz[b], z);
                        if (norm) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) z[q] = (z[q] - mean[q]) * rstd[q];
                        }
                        if (act == SSCG_ACT_RELU) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) gv[q] = z[q] > 0.f ? gv[q] : 0.f;
                        } else if (act == SSCG_ACT_LRELU) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) gv[q] = z[q] > 0.f ? gv[q] : gv[q] * a.slope;
                        } else if (act == SSCG_ACT_TANH) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) gv[q] = gv[q] * (1.f - z[q] * z[q]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) z[q] = 0.f;
                    }
                    *reinterpret_cast<uint4*>(dz0 + so) = pack8(gv);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        acc1[q] += gv[q];
                        acc2[q] += gv[q] * z[q];
                    }
                }
            }
        };
        if (border_unit) body(std::true_type{});
        else body(std::false_type{});
        ring.release(k);
    }
    if (a.bstats != nullptr) stream_flush_stats<kStrThreads>(s_red, acc1, acc2, a.bstats, cur_n, a.C, g.CH, chunk);
}

// ---------------------------------------------------------------------------------------------
// backward, second half: dRaw = rstd * (dZ - mean(dZ) - Z * mean(dZ * Z))
// ---------------------------------------------------------------------------------------------
template <int kStrThreads>
__global__ void __launch_bounds__(kStrThreads + 32, 1) in_bwd_apply_stream_kernel(const __grid_constant__ BwdStreamDev p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float2 s_pairs[kMaxStreamC];
    const SscgBwdArgs& a = p.a;
    const StreamGeom& g = p.g;
    const StreamRing ring = stream_ring_init<kStrThreads>(smem_raw, g);
    const int u0 = (int)((long long)blockIdx.x * g.total / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * g.total / gridDim.x);
    const int cnt = u1 - u0;
    if (cnt <= 0) return;
    if (threadIdx.x >= kStrThreads) {
        if (threadIdx.x == kStrThreads) {
            const uint8_t* rawp = reinterpret_cast<const uint8_t*>(a.raw);
            const uint8_t* dzp = reinterpret_cast<const uint8_t*>(a.dz);
            for (int k = 0; k < cnt; ++k) {
                const int u = u0 + k;          // units are contiguous in both tensors: unit u starts at u * ub bytes
                ring.acquire(k);
                const uint32_t bar = ring.full(k), dst = ring.stage(k);
                mbar_arrive_expect_tx(bar, (uint32_t)(2 * g.ub));
                bulk_load(dst, rawp + (long long)u * g.ub, g.ub, bar);
                bulk_load(dst + g.ub, dzp + (long long)u * g.ub, g.ub, bar);
            }
        }
        return;
    }
    const int chunk = threadIdx.x % g.CH;
    const int c0 = chunk * 8;
    const int prow = threadIdx.x / g.CH;
    const int pstep = kStrThreads / g.CH;
    const float inv_cnt = 1.f / (float)(a.H * a.W);
    float mean[8], rstd[8], m1[8], m2[8];
    int cur_n = -1;
    for (int k = 0; k < cnt; ++k) {
        const int u = u0 + k;
        const int n = u / g.ups;
        if (n != cur_n) {
            cur_n = n;
            cta_load_sums<false>(a.stats, n, a.C, inv_cnt, a.eps, s_pairs, threadIdx.x, kStrThreads, c0, true,
                                 BarNamed1<kStrThreads>(), mean, rstd);
            cta_load_sums<true>(a.bstats, n, a.C, inv_cnt, 0.f, s_pairs, threadIdx.x, kStrThreads, c0, true,
                                BarNamed1<kStrThreads>(), m1, m2);
        }
        mbar_wait(ring.full(k), ring.parity(k), 13);
        const uint32_t st = ring.stage(k);
        __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.draw) + (long long)u * (g.ub / 2) + c0;
        if (a.draw_pad > 0) {          // dRaw goes into a zero-haloed buffer (input of the N-expanded data gradient)
            const UnitPos up = unit_pos(g, u);
            obase = reinterpret_cast<__nv_bfloat16*>(p.draw) +
                    (((long long)n * (a.H + 2 * a.draw_pad) + up.h + a.draw_pad) * (a.W + 2 * a.draw_pad) + up.w0 + a.draw_pad) * a.C + c0;
        }
        const uint32_t pxb = (uint32_t)a.C * 2u;
        const uint32_t s_z0 = st + (uint32_t)c0 * 2u;
        uint8_t* ob = reinterpret_cast<uint8_t*>(obase);
        for (int px0 = prow; px0 < g.seg_px; px0 += 2 * pstep) {
            uint4 rz[2], rg[2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int px = px0 + b * pstep;
                if (px < g.seg_px) {
                    const uint32_t so = s_z0 + (uint32_t)px * pxb;
                    rz[b] = lds128(so);
                    rg[b] = lds128(so + g.ub);
                }
            }
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int px = px0 + b * pstep;
                if (px >= g.seg_px) continue;
                float z[8], gv[8];
                cvt8(rz[b], z);
                cvt8(rg[b], gv);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float zz = (z[q] - mean[q]) * rstd[q];
                    gv[q] = rstd[q] * (gv[q] - m1[q] - zz * m2[q]);
                }
                *reinterpret_cast<uint4*>(ob + (uint32_t)px * pxb) = pack8(gv);
            }
        }
        ring.release(k);
    }
}
