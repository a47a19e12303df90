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
#ifndef SSCG_STR_HEAVY
#define SSCG_STR_HEAVY 448
#endif
constexpr int kStrThreadsHeavy = SSCG_STR_HEAVY;
// Register budget: __launch_bounds__(544) is rounded up to 640 threads by ptxas (96 registers); __maxnreg__ states the
// real bound of 17 warps (65536 / 544 = 120).
#ifdef SSCG_STR_MAXNREG
#define SSCG_STR_BOUNDS(T) __maxnreg__(((65536 / ((T) + 32)) / 16) * 16)   /* allocation is per warp in units of 512 registers */
#else
#define SSCG_STR_BOUNDS(T) __launch_bounds__((T) + 32, 1)
#endif
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
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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

// Packed bf16x2 arithmetic for the two steps that are exact on already-rounded values: ReLU (rounding is
// monotonic and keeps the sign, so max(round(x), 0) == round(max(x, 0))) and the dropout scale (x2 / x0 is a
// power-of-two scaling).  One instruction per channel PAIR instead of two to three per channel.
__device__ __forceinline__ uint32_t bf2_relu(uint32_t a) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(0u));
    return d;
}
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// bits: keep-bit q for channel q of the vector (drop_bits).  y holds bit k at positions k and k + 15, so a left
// shift by 14 - 2j puts the bits of channels 2j / 2j + 1 at the positions of bf16 2.0 (0x4000) in both halves.
__device__ __forceinline__ void drop_scale_packed(uint4& v, uint32_t bits) {
    const uint32_t y = bits * 0x8001u;
    v.x = bf2_mul(v.x, (y << 14) & 0x40004000u);
    v.y = bf2_mul(v.y, (y << 12) & 0x40004000u);
    v.z = bf2_mul(v.z, (y << 10) & 0x40004000u);
    v.w = bf2_mul(v.w, (y << 8) & 0x40004000u);
}

// Ring of stages: full[s] completes when the bulk copies of a unit have landed, empty[s] when every
// consumer warp has finished reading it.  Warp kStrWarps is the producer; consumer warps run freely.
// Stage index and phase parity are carried by the unit walker below (no division per unit).
//
// Next to the barriers every stage has a 32-byte DESCRIPTOR written by the producer thread before it arms the stage:
// position of the unit and the byte offsets derived from it.  These values are uniform over the CTA; when each
// consumer thread derived them itself (mirror rows, three or four 64-bit multiply-add chains) they were ~140 of the
// ~450 instructions a warp spends per unit of a residual-block shape.
struct StreamRing {
    uint32_t base;          // shared-space address of stage 0
    uint32_t bars;          // shared-space address of full[0]; empty[s] follows the full barriers
    uint32_t descs;         // shared-space address of the stage descriptors (32 bytes each)
    int nst, stage_bytes;
    __device__ __forceinline__ uint32_t desc(int st) const { return descs + st * 32; }
    __device__ __forceinline__ uint32_t stage(int st) const { return base + st * stage_bytes; }
    __device__ __forceinline__ uint32_t full(int st) const { return bars + st * 8; }
    __device__ __forceinline__ uint32_t empty(int st) const { return bars + (kStrMaxStages + st) * 8; }
    // producer side: the stage is free once the consumers released its previous occupant (one ring round earlier)
    __device__ __forceinline__ void acquire(int st, uint32_t par, bool first_round) const {
        if (!first_round) mbar_wait(empty(st), par ^ 1u, 14);
    }
    // consumer side: all lanes of a warp are done with the stage
    __device__ __forceinline__ void release(int st) const {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty(st));
    }
};

template <int kStrThreads>
__device__ __forceinline__ StreamRing stream_ring_init(uint8_t* smem_raw, const StreamGeom& g) {
    constexpr int kStrWarps = kStrThreads / 32;
    StreamRing r;
    const uint32_t al = (smem_u32(smem_raw) + 127u) & ~127u;
    r.bars = al;
    r.descs = al + 128;
    r.base = al + 128 + kStrMaxStages * 32;
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

// Position of a CTA's current unit and its ring stage, advanced incrementally: the first version divided
// (u / ups, r / upr, k % nst, k / nst) per unit and thread, ~100 of the ~190 per-unit instructions of kernels
// whose consumers are issue-bound (the ring is always full when they arrive).
struct UnitWalk {
    int n, h, part, w0;     // sample, image row, unit inside the row, first pixel of the unit
    int st;                 // ring stage
    uint32_t par;           // phase parity of that stage's barriers
    bool first_round;
    __device__ __forceinline__ UnitWalk(const StreamGeom& g, int u) {
        n = u / g.ups;
        const int r = u - n * g.ups;
        h = r / g.upr;
        part = r - h * g.upr;
        w0 = part * g.seg_px;
        st = 0;
        par = 0;
        first_round = true;
    }
    __device__ __forceinline__ void next(const StreamGeom& g, int H) {
        w0 += g.seg_px;
        if (++part == g.upr) {
            part = 0;
            w0 = 0;
            if (++h == H) {
                h = 0;
                ++n;
            }
        }
        if (++st == g.nst) {
            st = 0;
            par ^= 1u;
            first_round = false;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// forward: y = dropout(act(instance_norm(raw))) (+ residual), written with (reflect) halo
// ---------------------------------------------------------------------------------------------
struct ApplyStreamDev {
    SscgApplyArgs a;
    StreamGeom g;
};

// SPEC folds the flags of the three shapes that make up a generator pass at compile time (the generic kernel spends
// about a third of its instructions on uniform flag tests, fp32 ReLU / dropout and their predication):
//   0  every flag read at run time
//   1  norm + ReLU                      (stem, down / up-sampling layers)
//   2  norm + ReLU + dropout            (first conv of a residual block)
//   3  norm + residual add, no act      (second conv of a residual block)
template <int kStrThreads, int SPEC>
__global__ void SSCG_STR_BOUNDS(kStrThreads) in_apply_stream_kernel(const __grid_constant__ ApplyStreamDev p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float2 s_pairs[kMaxStreamC];
    const SscgApplyArgs& a = p.a;
    const StreamGeom& g = p.g;
    const StreamRing ring = stream_ring_init<kStrThreads>(smem_raw, g);
    pdl_wait();          // everything above overlaps the predecessor's tail (sscg_common.cuh, launch_k)
    pdl_launch();
    const int u0 = (int)((long long)blockIdx.x * g.total / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * g.total / gridDim.x);
    const int cnt = u1 - u0;
    if (cnt <= 0) return;
    const bool norm = SPEC != 0 ? true : (a.stats != nullptr);
    const bool has_res = SPEC != 0 ? (SPEC == 3) : (a.res.ptr != nullptr);
    if (threadIdx.x >= kStrThreads) {
        // ================================ producer warp ========================================
        if (threadIdx.x == kStrThreads) {
            const __nv_bfloat16* rawp = reinterpret_cast<const __nv_bfloat16*>(a.raw);
            const __nv_bfloat16* resp = reinterpret_cast<const __nv_bfloat16*>(a.res.ptr);
            const bool reflect_p = a.pad > 0 && a.pad_mode == SSCG_PAD_REFLECT;
            const int Hp_p = a.H + 2 * a.pad, Wp_p = a.W + 2 * a.pad;
            UnitWalk uw(g, u0);
            for (int k = 0; k < cnt; ++k, uw.next(g, a.H)) {
                ring.acquire(uw.st, uw.par, uw.first_round);
                const uint32_t bar = ring.full(uw.st), dst = ring.stage(uw.st);
                // descriptor: {n, h, w0, hm1 | hm2, byte offset of the unit in dst, dropout index, -}
                int hm1 = -1, hm2 = -1;
                if (reflect_p) mirror_pos(uw.h, a.H, a.pad, hm1, hm2);
                const long long spix = ((long long)uw.n * a.H + uw.h) * a.W + uw.w0;
                const long long dpix = ((long long)uw.n * Hp_p + (uw.h + a.pad)) * Wp_p + (uw.w0 + a.pad);
                sts128(ring.desc(uw.st), (uint32_t)uw.n, (uint32_t)uw.h, (uint32_t)uw.w0, (uint32_t)hm1);
                sts128(ring.desc(uw.st) + 16, (uint32_t)hm2, (uint32_t)(dpix * a.C * 2), (uint32_t)(spix * g.CH), 0u);
                mbar_arrive_expect_tx(bar, (uint32_t)(g.ntens * g.ub));
                bulk_load(dst, rawp + spix * a.C, g.ub, bar);
                if (has_res)
                    bulk_load(dst + g.ub, resp + (long long)uw.n * a.res.sN + (long long)uw.h * a.res.sH +
                                              (long long)uw.w0 * a.res.sW, g.ub, bar);
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
    const uint64_t seed = (SPEC == 1 || SPEC == 3) ? 0ull
                          : ((a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                              : a.drop_seed);
    const bool drop = SPEC == 2 ? true : (seed != 0);
    const uint32_t pxb = (uint32_t)a.C * 2u;        // bytes per pixel
    float mean[8], rstd[8];
    int cur_n = -1;
    int st_i = 0;
    uint32_t par = 0;
    for (int k = 0; k < cnt; ++k) {
        mbar_wait(ring.full(st_i), par, 11);
        const uint4 ds0 = lds128(ring.desc(st_i)), ds1 = lds128(ring.desc(st_i) + 16);
        const int n = (int)ds0.x, h = (int)ds0.y, w0 = (int)ds0.z, hm1 = (int)ds0.w, hm2 = (int)ds1.x;
        if (n != cur_n) {
            cur_n = n;
            if (norm) cta_load_sums<false>(a.stats, n, a.C, 1.f / (float)(a.H * a.W), a.eps, s_pairs, threadIdx.x, kStrThreads,
                                           c0, true, BarNamed1<kStrThreads>(), mean, rstd);
        }
        const bool hmirror = (hm1 >= 0) || (hm2 >= 0);
        uint8_t* d0 = reinterpret_cast<uint8_t*>(a.dst) + ds1.y + c0 * 2;
        const uint32_t vec0 = ds1.z + (uint32_t)chunk;
        // does this unit hold a pixel that is mirrored into the halo?  (uniform per unit)
        const bool border_unit = reflect && (hmirror || w0 <= a.pad || w0 + g.seg_px - 1 >= a.W - 1 - a.pad);
        const uint32_t s_v0 = ring.stage(st_i) + (uint32_t)c0 * 2u;
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
                    uint4 out;
                    if constexpr (SPEC == 1 || SPEC == 2) {
                        out = pack8(v);
                        out.x = bf2_relu(out.x); out.y = bf2_relu(out.y); out.z = bf2_relu(out.z); out.w = bf2_relu(out.w);
                        if constexpr (SPEC == 2) drop_scale_packed(out, drop_bits(seed, vec0 + (uint32_t)px * (uint32_t)g.CH));
                    } else {
                        if constexpr (SPEC == 0) {
                            if (a.act == SSCG_ACT_RELU) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
                            } else if (a.act == SSCG_ACT_LRELU) {
#pragma unroll
                                for (int q = 0; q < 8; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * a.slope;
                            }
                            if (drop) {
                                const uint32_t bits = drop_bits(seed, vec0 + (uint32_t)px * (uint32_t)g.CH);
                                // keep -> x2, drop -> x0: bit q moved to the exponent position of 2.0f (0x40000000)
#pragma unroll
                                for (int q = 0; q < 8; ++q) v[q] *= __uint_as_float((bits << (30 - q)) & 0x40000000u);
                            }
                        }
                        if (has_res) {
                            float r8[8];
                            cvt8(rr[b], r8);
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] += r8[q];
                        }
                        out = pack8(v);
                    }
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
        ring.release(st_i);
        if (++st_i == g.nst) {
            st_i = 0;
            par ^= 1u;
        }
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
//
// In the specialised kernels z = (raw - mean) * rstd enters only through its sign (ReLU) and through the plane sum
// of dZ * z, so they keep d = raw - mean (sign(d) == sign(z): rstd > 0) and scale the sum by rstd once per flush.
template <int kStrThreads, int SPEC>
__global__ void SSCG_STR_BOUNDS(kStrThreads) in_bwd_prep_stream_kernel(const __grid_constant__ BwdStreamDev p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(16) float s_red[kStrThreads * 16];
    __shared__ float2 s_pairs[kMaxStreamC];
    const SscgBwdArgs& a = p.a;
    const StreamGeom& g = p.g;
    const StreamRing ring = stream_ring_init<kStrThreads>(smem_raw, g);
    pdl_wait();          // everything above overlaps the predecessor's tail (sscg_common.cuh, launch_k)
    pdl_launch();
    const int u0 = (int)((long long)blockIdx.x * g.total / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * g.total / gridDim.x);
    const int cnt = u1 - u0;
    if (cnt <= 0) return;
    const bool norm = SPEC != 0 ? true : (a.stats != nullptr);
    const int act = SPEC == 1 ? SSCG_ACT_RELU : (SPEC == 2 ? SSCG_ACT_NONE : a.act);
    const bool need_raw = norm || act != SSCG_ACT_NONE;
    const bool has_skip = SPEC == 1 ? false : (SPEC == 2 ? true : (a.skip.ptr != nullptr));
    const bool has_gout = SPEC == 1 ? false : (SPEC == 2 ? true : (a.g_out != nullptr));
    // without activation and dropout dZ IS the folded total gradient: the caller may pass the same buffer for both
    // (second conv of a residual block) and the value is stored once
    const bool store_gout = has_gout && a.g_out != a.dz;
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
            UnitWalk up(g, u0);
            for (int k = 0; k < cnt; ++k, up.next(g, a.H)) {
                ring.acquire(up.st, up.par, up.first_round);
                const uint32_t bar = ring.full(up.st), dst = ring.stage(up.st);
                const int lpad = (fold && up.w0 == 0) ? a.pad : 0;
                const int rpad = (fold && up.w0 == last_part_w0) ? a.pad : 0;
                const uint32_t dy_bytes = (uint32_t)((g.seg_px + lpad + rpad) * a.C * 2);
                // descriptor: {n, h, w0, hm1 | hm2, byte offset of the unit in dz, in g_out / raw, dropout index}
                int hm1 = -1, hm2 = -1;
                if (fold) mirror_pos(up.h, a.H, a.pad, hm1, hm2);
                const long long spix = ((long long)up.n * a.H + up.h) * a.W + up.w0;
                const long long dzpix = a.dz_pad > 0 ? (((long long)up.n * (a.H + 2 * a.dz_pad) + up.h + a.dz_pad) * (a.W + 2 * a.dz_pad) +
                                                        up.w0 + a.dz_pad)
                                                     : spix;
                sts128(ring.desc(up.st), (uint32_t)up.n, (uint32_t)up.h, (uint32_t)up.w0, (uint32_t)hm1);
                sts128(ring.desc(up.st) + 16, (uint32_t)hm2, (uint32_t)(dzpix * a.C * 2), (uint32_t)(spix * a.C * 2),
                       (uint32_t)(spix * g.CH));
                mbar_arrive_expect_tx(bar, dy_bytes + (uint32_t)((g.ntens - 1) * g.ub));
                bulk_load(dst, dyp + (long long)up.n * a.dyp.sN + (long long)(up.h + a.pad) * a.dyp.sH +
                                   (long long)(up.w0 + a.pad - lpad) * a.dyp.sW, dy_bytes, bar);
                if (has_skip)
                    bulk_load(dst + off_skip, skp + (long long)up.n * a.skip.sN + (long long)up.h * a.skip.sH +
                                                  (long long)up.w0 * a.skip.sW, g.ub, bar);
                if (need_raw) bulk_load(dst + off_raw, rawp + spix * a.C, g.ub, bar);
            }
        }
        return;
    }
    const int chunk = threadIdx.x % g.CH;
    const int c0 = chunk * 8;
    const int prow = threadIdx.x / g.CH;
    const int pstep = kStrThreads / g.CH;
    // seg_px is not a multiple of pstep for most shapes (32 pixels over 14 rows of threads): the rows that get the
    // extra pixel rotate from unit to unit, otherwise the same four warps would set the pace of every unit
    const int rot_step = g.seg_px % pstep;
    int pfirst = prow;
    const uint64_t seed = SPEC == 2 ? 0ull
                          : ((a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                              : a.drop_seed);
    const bool drop = seed != 0;
    const uint32_t pxb = (uint32_t)a.C * 2u;                                   // bytes per pixel
    float mean[8], rstd[8], acc1[8], acc2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc1[q] = acc2[q] = 0.f;
    auto flush = [&](int n) {
        if (SPEC != 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc2[q] *= s_pairs[c0 + q].y;          // rstd of sample n (still in s_pairs)
        }
        stream_flush_stats<kStrThreads>(s_red, acc1, acc2, a.bstats, n, a.C, g.CH, chunk);
    };
    int cur_n = -1;
    int st_i = 0;
    uint32_t par = 0;
    for (int k = 0; k < cnt; ++k) {
        mbar_wait(ring.full(st_i), par, 12);
        const uint32_t st = ring.stage(st_i);
        const uint4 d0 = lds128(ring.desc(st_i)), d1 = lds128(ring.desc(st_i) + 16);
        const int n = (int)d0.x, h = (int)d0.y, w0 = (int)d0.z, hm1 = (int)d0.w, hm2 = (int)d1.x;
        if (n != cur_n) {
            if (cur_n >= 0 && a.bstats != nullptr) flush(cur_n);
            cur_n = n;
            if (norm) cta_load_sums<false>(a.stats, n, a.C, 1.f / (float)(a.H * a.W), a.eps, s_pairs, threadIdx.x, kStrThreads,
                                           c0, true, BarNamed1<kStrThreads>(), mean, rstd);
        }
        const bool hborder = (hm1 >= 0) || (hm2 >= 0);
        const int lpad = (fold && w0 == 0) ? a.pad : 0;
        const int col0 = w0 + a.pad - lpad;            // padded column held by staged pixel 0 of the gradient row
        const uint32_t s_g0 = st + (uint32_t)lpad * pxb + (uint32_t)c0 * 2u;      // gradient of pixel w0
        const uint32_t s_k0 = st + (uint32_t)off_skip + (uint32_t)c0 * 2u;
        const uint32_t s_z0 = st + (uint32_t)off_raw + (uint32_t)c0 * 2u;
        uint8_t* gout0 = reinterpret_cast<uint8_t*>(a.g_out) + d1.z + c0 * 2;
        uint8_t* dz0 = reinterpret_cast<uint8_t*>(a.dz) + d1.y + c0 * 2;
        const uint32_t vec0 = d1.w + (uint32_t)chunk;                               // dropout hash index of pixel w0 (mod 2^32)
        // rows next to the top / bottom border: the gradient of the mirrored halo row (same column) is fetched
        // from global memory up front, together with the shared-memory loads of the batch
        const int xf = hm1 >= 0 ? 1 : 2;
        const uint8_t* mrow0 = nullptr;
        if (hborder)
            mrow0 = reinterpret_cast<const uint8_t*>(
                reinterpret_cast<const __nv_bfloat16*>(a.dyp.ptr) + (long long)n * a.dyp.sN +
                (long long)(hm1 >= 0 ? hm1 : hm2) * a.dyp.sH + (long long)(w0 + a.pad) * a.dyp.sW + c0);
        // does this unit hold a pixel whose gradient receives mirrored halo contributions?  (uniform per unit)
        const bool border_unit = hborder || (fold && (w0 <= a.pad || w0 + g.seg_px - 1 >= a.W - 1 - a.pad));
        auto body = [&](auto border_tag) {
            constexpr bool BORDER = decltype(border_tag)::value;
            for (int px0 = pfirst; px0 < g.seg_px; px0 += 2 * pstep) {
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
                    const int w = w0 + px;
                    const bool bpix = BORDER && (hborder || (w <= a.pad || w >= a.W - 1 - a.pad));
                    // dropout: x2 / x0 is exact on the bf16 value itself as long as nothing has been added to it (and the
                    // pre-dropout total gradient is not an output)
                    uint32_t bits = 0;
                    if (drop) bits = drop_bits(seed, vec0 + (uint32_t)px * (uint32_t)g.CH);
                    const bool packed_drop = drop && !has_skip && !has_gout && !bpix;
                    if (packed_drop) drop_scale_packed(rg[b], bits);
                    cvt8(rg[b], gv);
                    if (BORDER) {
                        if (bpix) {
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
                    if (store_gout) *reinterpret_cast<uint4*>(gout0 + so) = pack8(gv);
                    if (drop && !packed_drop) {
                        // keep -> x2, drop -> x0: bit q moved to the exponent position of 2.0f (0x40000000)
#pragma unroll
                        for (int q = 0; q < 8; ++q) gv[q] *= __uint_as_float((bits << (30 - q)) & 0x40000000u);
                    }
                    if constexpr (SPEC != 0) {
                        cvt8(rz[b], z);
#pragma unroll
                        for (int q = 0; q < 8; ++q) z[q] -= mean[q];
                        if constexpr (SPEC == 1) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) gv[q] = z[q] > 0.f ? gv[q] : 0.f;
                        }
                    } else if (need_raw) {
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
        ring.release(st_i);
        if (++st_i == g.nst) {
            st_i = 0;
            par ^= 1u;
        }
        pfirst -= rot_step;
        if (pfirst < 0) pfirst += pstep;
    }
    if (a.bstats != nullptr) flush(cur_n);
}

// ---------------------------------------------------------------------------------------------
// backward, second half: dRaw = rstd * (dZ - mean(dZ) - Z * mean(dZ * Z))
// ---------------------------------------------------------------------------------------------
template <int kStrThreads>
__global__ void SSCG_STR_BOUNDS(kStrThreads) in_bwd_apply_stream_kernel(const __grid_constant__ BwdStreamDev p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ float2 s_pairs[kMaxStreamC];
    const SscgBwdArgs& a = p.a;
    const StreamGeom& g = p.g;
    const StreamRing ring = stream_ring_init<kStrThreads>(smem_raw, g);
    pdl_wait();          // everything above overlaps the predecessor's tail (sscg_common.cuh, launch_k)
    pdl_launch();
    const int u0 = (int)((long long)blockIdx.x * g.total / gridDim.x);
    const int u1 = (int)((long long)(blockIdx.x + 1) * g.total / gridDim.x);
    const int cnt = u1 - u0;
    if (cnt <= 0) return;
    if (threadIdx.x >= kStrThreads) {
        if (threadIdx.x == kStrThreads) {
            const uint8_t* rawp = reinterpret_cast<const uint8_t*>(a.raw) + (long long)u0 * g.ub;
            const uint8_t* dzp = reinterpret_cast<const uint8_t*>(a.dz) + (long long)u0 * g.ub;
            UnitWalk uw(g, u0);
            for (int k = 0; k < cnt; ++k, uw.next(g, a.H)) {     // units are contiguous in both tensors
                ring.acquire(uw.st, uw.par, uw.first_round);
                const uint32_t bar = ring.full(uw.st), dst = ring.stage(uw.st);
                mbar_arrive_expect_tx(bar, (uint32_t)(2 * g.ub));
                bulk_load(dst, rawp, g.ub, bar);
                bulk_load(dst + g.ub, dzp, g.ub, bar);
                rawp += g.ub;
                dzp += g.ub;
            }
        }
        return;
    }
    const int chunk = threadIdx.x % g.CH;
    const int c0 = chunk * 8;
    const int prow = threadIdx.x / g.CH;
    const int pstep = kStrThreads / g.CH;
    const float inv_cnt = 1.f / (float)(a.H * a.W);
    const uint32_t pxb = (uint32_t)a.C * 2u;
    // dRaw = rstd * (dZ - m1 - zhat * m2) with zhat = (z - mean) * rstd, as two FMAs per element:
    //   dRaw = ca * dZ + (cb * z + cc),  ca = rstd, cb = -rstd^2 * m2, cc = -rstd * m1 - mean * cb
    // (five dependent operations in the literal form; this kernel is issue-bound).  The constants differ from the
    // literal form by fp32 rounding only, far below the bf16 rounding of the result.
    float ca[8], cb[8], cc[8];
    int cur_n = -1;
    UnitWalk uw(g, u0);
    for (int k = 0; k < cnt; ++k, uw.next(g, a.H)) {
        const int n = uw.n;
        if (n != cur_n) {
            cur_n = n;
            float mean[8], m1[8];
            cta_load_sums<false>(a.stats, n, a.C, inv_cnt, a.eps, s_pairs, threadIdx.x, kStrThreads, c0, true,
                                 BarNamed1<kStrThreads>(), mean, ca);
            cta_load_sums<true>(a.bstats, n, a.C, inv_cnt, 0.f, s_pairs, threadIdx.x, kStrThreads, c0, true,
                                BarNamed1<kStrThreads>(), m1, cb);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                cb[q] = -ca[q] * ca[q] * cb[q];
                cc[q] = -ca[q] * m1[q] - mean[q] * cb[q];
            }
        }
        uint8_t* ob;
        if (a.draw_pad > 0)            // dRaw goes into a zero-haloed buffer (input of the N-expanded data gradient)
            ob = reinterpret_cast<uint8_t*>(p.draw) +
                 ((((long long)n * (a.H + 2 * a.draw_pad) + uw.h + a.draw_pad) * (a.W + 2 * a.draw_pad) + uw.w0 + a.draw_pad) * a.C + c0) * 2;
        else
            ob = reinterpret_cast<uint8_t*>(p.draw) + (long long)(u0 + k) * g.ub + c0 * 2;
        mbar_wait(ring.full(uw.st), uw.par, 13);
        const uint32_t s_z0 = ring.stage(uw.st) + (uint32_t)c0 * 2u;
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
                for (int q = 0; q < 8; ++q) gv[q] = fmaf(ca[q], gv[q], fmaf(cb[q], z[q], cc[q]));
                *reinterpret_cast<uint4*>(ob + (uint32_t)px * pxb) = pack8(gv);
            }
        }
        ring.release(uw.st);
    }
}
