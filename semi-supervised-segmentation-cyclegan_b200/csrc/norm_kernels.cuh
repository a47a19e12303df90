// norm_kernels.cuh — InstanceNorm apply / backward kernels (included by elementwise.cu inside
// namespace sscg, after its load/store helpers).
//
// All three are streaming kernels over NHWC planes: a thread owns one 8-channel vector (16 B of
// bf16) and walks pixels.  They are HBM/L2-bandwidth kernels, so each thread keeps a batch of
// independent 16-byte loads in flight before it touches any of them — with one load in flight per
// thread the SMs cannot cover the ~1 us memory latency (2.5 TB/s measured).  The loads are kept as
// packed 128-bit registers until they are consumed, so that the batch does not cost occupancy.
// Template parameter ANYF32: false = every tensor is bf16 (fast mode, 4 registers per vector in
// flight); true = per-tensor runtime dtype flags (bf16x3 parity mode and mixed cases).
#pragma once

template <bool ANYF32>
struct RawVec;
template <>
struct RawVec<false> {
    uint4 a;
};
template <>
struct RawVec<true> {
    uint4 a, b;
};

template <bool ANYF32>
__device__ __forceinline__ void raw_load(RawVec<ANYF32>& r, const void* base, bool fp32, long long off) {
    if constexpr (ANYF32) {
        if (fp32) {
            const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + off);
            r.a = p[0];
            r.b = p[1];
            return;
        }
    }
    r.a = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
}
template <bool ANYF32>
__device__ __forceinline__ void raw_cvt(const RawVec<ANYF32>& r, bool fp32, float (&v)[8]) {
    if constexpr (ANYF32) {
        if (fp32) {
            v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y);
            v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
            v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y);
            v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
            return;
        }
    }
    const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[2 * q] = __uint_as_float(w[q] << 16);
        v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
    }
}

// Running (row, column) of a thread's pixel sequence pix0, pix0 + step, pix0 + 2*step, ...: one
// div/mod at the start, then additions only (the per-pixel 64-bit index arithmetic and divisions
// were a third of the instructions of these kernels).
struct PixWalk {
    int pix, h, w;
    __device__ __forceinline__ PixWalk(int pix0, int W) : pix(pix0), h(pix0 / W), w(pix0 % W) {}
    __device__ __forceinline__ void advance(int step, int W) {
        pix += step;
        w += step;
        while (w >= W) { w -= W; ++h; }
    }
};

// ---------------------------------------------------------------------------------------------
// forward: y = dropout(act(instance_norm(raw))) (+ residual), written with halo
// ---------------------------------------------------------------------------------------------
struct ApplyDev {
    SscgApplyArgs a;
    int CH;        // 8-channel vectors per pixel
    int rows;      // pixels per pass per block
    int iters;     // passes per block (multiple of the batch)
};

#ifndef SSCG_APPLY_BATCH
#define SSCG_APPLY_BATCH 4
#endif
#ifndef SSCG_APPLY_MINB
#define SSCG_APPLY_MINB 3
#endif
constexpr int kApplyBatch = SSCG_APPLY_BATCH;

template <bool ANYF32>
__global__ void __launch_bounds__(256, SSCG_APPLY_MINB) in_apply_kernel(const __grid_constant__ ApplyDev p) {
    pdl_wait();
    pdl_launch();
    const SscgApplyArgs& a = p.a;
    const int n = blockIdx.y;
    const int chunk = threadIdx.x % p.CH;
    const int row = threadIdx.x / p.CH;
    const int Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad;
    const int c0 = chunk * 8;
    float mean[8], rstd[8];
    const bool norm = a.stats != nullptr;
    __shared__ float2 s_pairs[kMaxNormC];
    if (norm) cta_load_sums<false>(a.stats, n, a.C, 1.f / (float)(a.H * a.W), a.eps, s_pairs, threadIdx.x, 256, c0,
                                   row < p.rows, BarBlock(), mean, rstd);
    if (row >= p.rows) return;
    const bool has_res = a.res.ptr != nullptr;
    const bool res_lo = ANYF32 && a.res_lo != nullptr;
    const bool raw_f32 = ANYF32 && a.raw_fp32 != 0;
    void* dst_lo = ANYF32 ? a.dst_lo : nullptr;
    const uint64_t seed = (a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                           : a.drop_seed;
    const int npix = Hp * Wp;
    PixWalk walk(blockIdx.x * p.rows * p.iters + row, Wp);      // walks the PADDED destination pixels
    const long long dbase = (long long)n * npix * a.C + c0;
    const long long sbase = (long long)n * a.H * a.W;
    for (int it0 = 0; it0 < p.iters; it0 += kApplyBatch) {
        RawVec<ANYF32> rv[kApplyBatch];
        RawVec<false> rr[kApplyBatch], rl[kApplyBatch];
        int spix[kApplyBatch], dpix[kApplyBatch];
        int state[kApplyBatch];   // 0: skip, 1: zero halo, 2: data
        // ---- issue all loads of the batch ----------------------------------------------------
#pragma unroll
        for (int b = 0; b < kApplyBatch; ++b) {
            state[b] = 0;
            spix[b] = 0;
            dpix[b] = walk.pix;
            if (walk.pix < npix) {
                int h = walk.h - a.pad, w = walk.w - a.pad;
                bool inside = true;
                if (a.pad_mode == SSCG_PAD_REFLECT) {
                    h = reflect_idx(h, a.H);
                    w = reflect_idx(w, a.W);
                } else {
                    inside = !(h < 0 || h >= a.H || w < 0 || w >= a.W);
                }
                state[b] = inside ? 2 : 1;
                if (inside) {
                    spix[b] = h * a.W + w;
                    raw_load<ANYF32>(rv[b], a.raw, raw_f32, (sbase + spix[b]) * a.C + c0);
                    if (has_res) {
                        const long long ro = (long long)n * a.res.sN + (long long)h * a.res.sH + (long long)w * a.res.sW + c0;
                        raw_load<false>(rr[b], a.res.ptr, false, ro);
                        if (res_lo) raw_load<false>(rl[b], a.res_lo, false, ro);
                    }
                }
            }
            walk.advance(p.rows, Wp);
        }
        // ---- compute + store -------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < kApplyBatch; ++b) {
            if (state[b] == 0) continue;
            const long long doff = dbase + (long long)dpix[b] * a.C;
            float v[8];
            if (state[b] == 1) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
                store8_bf16(a.dst, dst_lo, doff, v);
                continue;
            }
            raw_cvt<ANYF32>(rv[b], raw_f32, v);
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
                const uint32_t bits = drop_bits(seed, (unsigned long long)(sbase + spix[b]) * p.CH + chunk);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = ((bits >> q) & 1u) ? 2.f * v[q] : 0.f;
            }
            if (has_res) {
                float r[8];
                raw_cvt<false>(rr[b], false, r);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] += r[q];
                if (res_lo) {
                    raw_cvt<false>(rl[b], false, r);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] += r[q];
                }
            }
            store8_bf16(a.dst, dst_lo, doff, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward: dZ and the two per-plane reductions, then dRaw
// ---------------------------------------------------------------------------------------------
struct BwdDev {
    SscgBwdArgs a;
    void* draw; void* draw_lo;
    int CH, rows, iters;
};


// positions of the padded gradient buffer that fold onto source index s (reflect) — at most 3
__device__ __forceinline__ int fold_positions(int s, int n, int pad, int mode, int (&q)[3]) {
    int cnt = 0;
    q[cnt++] = s + pad;
    if (mode == SSCG_PAD_REFLECT) {
        if (s >= 1 && s <= pad) q[cnt++] = pad - s;
        if (s <= n - 2 && s >= n - 1 - pad) q[cnt++] = pad + 2 * (n - 1) - s;
    }
    return cnt;
}

#ifndef SSCG_PREP_BATCH
#define SSCG_PREP_BATCH 4
#endif
#ifndef SSCG_PREP_MINB
#define SSCG_PREP_MINB 2
#endif
constexpr int kPrepBatch = SSCG_PREP_BATCH;

// First half of the backward: writes dZ (and the folded total gradient), produces the plane sums.
template <bool ANYF32>
__global__ void __launch_bounds__(256, SSCG_PREP_MINB) in_bwd_prep_kernel(const __grid_constant__ BwdDev p) {
    pdl_wait();
    pdl_launch();
    const SscgBwdArgs& a = p.a;
    __shared__ float s_red[256 * 16];
    const int n = blockIdx.y;
    const int chunk = threadIdx.x % p.CH;
    const int row = threadIdx.x / p.CH;
    const bool active = row < p.rows;
    const int c0 = chunk * 8;
    float mean[8], rstd[8];
    const bool norm = a.stats != nullptr;
    const bool need_raw = norm || a.act != SSCG_ACT_NONE;
    const bool fold = (a.pad_mode == SSCG_PAD_REFLECT) && a.pad > 0;
    const bool has_dyp = a.dyp.ptr != nullptr, has_skip = a.skip.ptr != nullptr;
    const bool dyp_f32 = ANYF32 && a.dyp_fp32 != 0, skip_f32 = ANYF32 && a.skip_fp32 != 0;
    const bool raw_f32 = ANYF32 && a.raw_fp32 != 0;
    __shared__ float2 s_pairs[kMaxNormC];
    if (norm) cta_load_sums<false>(a.stats, n, a.C, 1.f / (float)(a.H * a.W), a.eps, s_pairs, threadIdx.x, 256, c0, active,
                                   BarBlock(), mean, rstd);
    const uint64_t seed = (a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                           : a.drop_seed;
    float acc1[8], acc2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc1[q] = acc2[q] = 0.f;
    const int npix = a.H * a.W;
    auto sweep = [&]() {
        PixWalk walk(blockIdx.x * p.rows * p.iters + row, a.W);
        const long long obase = (long long)n * npix * a.C + c0;                 // raw / dz / g_out (unpadded NHWC)
        const long long ybase = (long long)n * a.dyp.sN + (long long)a.pad * a.dyp.sH + (long long)a.pad * a.dyp.sW + c0;
        const long long kbase = (long long)n * a.skip.sN + c0;
        for (int it0 = 0; it0 < p.iters; it0 += kPrepBatch) {
            RawVec<ANYF32> rg[kPrepBatch], rs[kPrepBatch], rz[kPrepBatch];
            int ph[kPrepBatch], pw[kPrepBatch], pp[kPrepBatch];
            // ---- loads (interior gradient position, skip gradient, raw activation) -------------
#pragma unroll
            for (int b = 0; b < kPrepBatch; ++b) {
                ph[b] = walk.h; pw[b] = walk.w; pp[b] = walk.pix;
                if (walk.pix < npix) {
                    if (has_dyp)
                        raw_load<ANYF32>(rg[b], a.dyp.ptr, dyp_f32,
                                         ybase + (long long)walk.h * a.dyp.sH + (long long)walk.w * a.dyp.sW);
                    if (has_skip)
                        raw_load<ANYF32>(rs[b], a.skip.ptr, skip_f32,
                                         kbase + (long long)walk.h * a.skip.sH + (long long)walk.w * a.skip.sW);
                    if (need_raw) raw_load<ANYF32>(rz[b], a.raw, raw_f32, obase + (long long)walk.pix * a.C);
                }
                walk.advance(p.rows, a.W);
            }
            // ---- compute + store ---------------------------------------------------------------
#pragma unroll
            for (int b = 0; b < kPrepBatch; ++b) {
                const int pix = pp[b];
                if (pix >= npix) continue;
                const long long spix = (long long)n * npix + pix;
                const long long off = obase + (long long)pix * a.C;
                float g[8], z[8];
                if (has_dyp) {
                    raw_cvt<ANYF32>(rg[b], dyp_f32, g);
                    const int w = pw[b], h = ph[b];
                    // halo positions that mirror onto this pixel exist only within `pad` of the border
                    if (fold && (h <= a.pad || w <= a.pad || h >= a.H - 1 - a.pad || w >= a.W - 1 - a.pad)) {
                        int hq[3], wq[3];
                        const int nh = fold_positions(h, a.H, a.pad, a.pad_mode, hq);
                        const int nw = fold_positions(w, a.W, a.pad, a.pad_mode, wq);
                        if (nh > 1 || nw > 1) {
                            for (int x = 0; x < nh; ++x)
                                for (int y = 0; y < nw; ++y) {
                                    if (x == 0 && y == 0) continue;
                                    float t[8];
                                    load8(a.dyp.ptr, dyp_f32,
                                          (long long)n * a.dyp.sN + (long long)hq[x] * a.dyp.sH + (long long)wq[y] * a.dyp.sW + c0, t);
#pragma unroll
                                    for (int q = 0; q < 8; ++q) g[q] += t[q];
                                }
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] = 0.f;
                }
                if (has_skip) {
                    float t[8];
                    raw_cvt<ANYF32>(rs[b], skip_f32, t);
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] += t[q];
                }
                if (a.g_out != nullptr) {
                    if (ANYF32 && a.g_fp32) store8_f32(a.g_out, off, g);
                    else store8_bf16(a.g_out, nullptr, off, g);
                }
                if (seed != 0) {
                    const uint32_t bits = drop_bits(seed, (unsigned long long)spix * p.CH + chunk);
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] = ((bits >> q) & 1u) ? 2.f * g[q] : 0.f;
                }
                if (need_raw) {
                    raw_cvt<ANYF32>(rz[b], raw_f32, z);
                    if (norm) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) z[q] = (z[q] - mean[q]) * rstd[q];
                    }
                    if (a.act == SSCG_ACT_RELU) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[q] = z[q] > 0.f ? g[q] : 0.f;
                    } else if (a.act == SSCG_ACT_LRELU) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[q] = z[q] > 0.f ? g[q] : g[q] * a.slope;
                    } else if (a.act == SSCG_ACT_TANH) {   // raw holds y = tanh(.)
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[q] = g[q] * (1.f - z[q] * z[q]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) z[q] = 0.f;
                }
                {
                    const long long doff = a.dz_pad > 0
                        ? ((((long long)n * (a.H + 2 * a.dz_pad) + ph[b] + a.dz_pad) * (a.W + 2 * a.dz_pad) + pw[b] + a.dz_pad) * a.C + c0)
                        : off;
                    if (ANYF32 && a.dz_fp32) store8_f32(a.dz, doff, g);
                    else store8_bf16(a.dz, ANYF32 ? a.dz_lo : nullptr, doff, g);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    acc1[q] += g[q];
                    acc2[q] += g[q] * z[q];
                }
            }
        }
    };
    if (active) sweep();
    if (a.bstats == nullptr) return;
    // block reduction over the pixel rows that share a channel vector: one smem pass, one sync
    {
        float* mine = s_red + threadIdx.x * 16;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            mine[2 * q] = active ? acc1[q] : 0.f;
            mine[2 * q + 1] = active ? acc2[q] : 0.f;
        }
    }
    __syncthreads();
    const int nout = p.CH * 16;    // (chunk, q, {sum, sum*z}) values of this block
    unsigned long long* bacc = reinterpret_cast<unsigned long long*>(a.bstats) + (long long)n * a.C * 2 * kDetWords;
    for (int o = threadIdx.x; o < nout; o += 256) {
        const int ch = o >> 4, e = o & 15;
        float s = 0.f;
        for (int r = 0; r < p.rows; ++r) s += s_red[(r * p.CH + ch) * 16 + e];
        // o = ch * 16 + 2 * q + k  ->  accumulator of bstats[n][ch * 8 + q][k]: order-independent integer sums (sscg_ptx.cuh)
        det_red_add(bacc + (long long)o * kDetWords, s);
    }
}

#ifndef SSCG_BAPPLY_BATCH
#define SSCG_BAPPLY_BATCH 4
#endif
#ifndef SSCG_BAPPLY_MINB
#define SSCG_BAPPLY_MINB 3
#endif
constexpr int kBwdApplyBatch = SSCG_BAPPLY_BATCH;

template <bool ANYF32>
__global__ void __launch_bounds__(256, SSCG_BAPPLY_MINB) in_bwd_apply_kernel(const __grid_constant__ BwdDev p) {
    pdl_wait();
    pdl_launch();
    const SscgBwdArgs& a = p.a;
    const int n = blockIdx.y;
    const int chunk = threadIdx.x % p.CH;
    const int row = threadIdx.x / p.CH;
    const int c0 = chunk * 8;
    const float inv_cnt = 1.f / (float)(a.H * a.W);
    const bool raw_f32 = ANYF32 && a.raw_fp32 != 0, dz_f32 = ANYF32 && a.dz_fp32 != 0;
    float mean[8], rstd[8], m1[8], m2[8];
    __shared__ float2 s_pairs[kMaxNormC];
    cta_load_sums<false>(a.stats, n, a.C, inv_cnt, a.eps, s_pairs, threadIdx.x, 256, c0, row < p.rows, BarBlock(), mean, rstd);
    cta_load_sums<true>(a.bstats, n, a.C, inv_cnt, 0.f, s_pairs, threadIdx.x, 256, c0, row < p.rows, BarBlock(), m1, m2);
    if (row >= p.rows) return;
    const int npix = a.H * a.W;
    const int pix0 = blockIdx.x * p.rows * p.iters + row;
    const long long obase = (long long)n * npix * a.C + c0;
    for (int it0 = 0; it0 < p.iters; it0 += kBwdApplyBatch) {
        RawVec<ANYF32> rz[kBwdApplyBatch], rg[kBwdApplyBatch];
#pragma unroll
        for (int b = 0; b < kBwdApplyBatch; ++b) {
            const int pix = pix0 + (it0 + b) * p.rows;
            if (pix < npix) {
                const long long off = obase + (long long)pix * a.C;
                raw_load<ANYF32>(rz[b], a.raw, raw_f32, off);
                raw_load<ANYF32>(rg[b], a.dz, dz_f32, off);
            }
        }
#pragma unroll
        for (int b = 0; b < kBwdApplyBatch; ++b) {
            const int pix = pix0 + (it0 + b) * p.rows;
            if (pix >= npix) continue;
            const long long off = obase + (long long)pix * a.C;
            float z[8], g[8];
            raw_cvt<ANYF32>(rz[b], raw_f32, z);
            raw_cvt<ANYF32>(rg[b], dz_f32, g);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float zz = (z[q] - mean[q]) * rstd[q];
                g[q] = rstd[q] * (g[q] - m1[q] - zz * m2[q]);
            }
            long long doff = off;
            if (a.draw_pad > 0) {
                const int h = pix / a.W, w = pix - h * a.W;
                doff = (((long long)n * (a.H + 2 * a.draw_pad) + h + a.draw_pad) * (a.W + 2 * a.draw_pad) + w + a.draw_pad) * a.C + c0;
            }
            store8_bf16(p.draw, ANYF32 ? p.draw_lo : nullptr, doff, g);
        }
    }
}
