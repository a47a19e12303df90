// norm_kernels.cuh — InstanceNorm apply / backward kernels (included by elementwise.cu).
//
// All three are streaming kernels over NHWC planes: a thread owns one 8-channel vector (16 B of
// bf16) and walks pixels.  They are HBM/L2-bandwidth kernels, so each thread keeps a batch of
// independent 16-byte loads in flight (kBatch pixels) before it touches any of them — with one
// load in flight per thread the SMs cannot cover the ~1 us memory latency and the kernels stall
// at ~2.5 TB/s (measured, profiles/r01_*).
#pragma once
// (included inside namespace sscg, after the load/store helpers of elementwise.cu)

// ---------------------------------------------------------------------------------------------
// forward: y = dropout(act(instance_norm(raw))) (+ residual), written with halo
// ---------------------------------------------------------------------------------------------
struct ApplyDev {
    SscgApplyArgs a;
    int CH;        // 8-channel vectors per pixel
    int rows;      // pixels per pass per block
    int iters;     // passes per block (multiple of kApplyBatch)
};

constexpr int kApplyBatch = 4;

__global__ void __launch_bounds__(256) in_apply_kernel(const __grid_constant__ ApplyDev p) {
    const SscgApplyArgs& a = p.a;
    const int n = blockIdx.y;
    const int chunk = threadIdx.x % p.CH;
    const int row = threadIdx.x / p.CH;
    if (row >= p.rows) return;
    const int Hp = a.H + 2 * a.pad, Wp = a.W + 2 * a.pad;
    const int c0 = chunk * 8;
    float mean[8], rstd[8];
    const bool norm = a.stats != nullptr;
    if (norm) load_norm(a.stats, a.eps, (long long)n * a.C + c0, 1.f / (float)(a.H * a.W), mean, rstd);
    const bool has_res = a.res.ptr != nullptr;
    const uint64_t seed = (a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                           : a.drop_seed;
    const long long npix = (long long)Hp * Wp;
    const long long pix0 = (long long)blockIdx.x * p.rows * p.iters + row;
    for (int it0 = 0; it0 < p.iters; it0 += kApplyBatch) {
        float v[kApplyBatch][8], r[kApplyBatch][8];
        long long spix[kApplyBatch], doff[kApplyBatch];
        int state[kApplyBatch];   // 0: skip, 1: zero halo, 2: data
        // ---- issue all loads of the batch ----------------------------------------------------
#pragma unroll
        for (int b = 0; b < kApplyBatch; ++b) {
            const long long pix = pix0 + (long long)(it0 + b) * p.rows;
            state[b] = 0;
            if (pix >= npix) continue;
            const int wp = pix % Wp, hp = pix / Wp;
            int h = hp - a.pad, w = wp - a.pad;
            doff[b] = (((long long)n * Hp + hp) * Wp + wp) * a.C + c0;
            if (a.pad_mode == SSCG_PAD_REFLECT) {
                h = reflect_idx(h, a.H);
                w = reflect_idx(w, a.W);
            } else if (h < 0 || h >= a.H || w < 0 || w >= a.W) {
                state[b] = 1;
                continue;
            }
            state[b] = 2;
            spix[b] = ((long long)n * a.H + h) * a.W + w;
            load8(a.raw, a.raw_fp32 != 0, spix[b] * a.C + c0, v[b]);
            if (has_res)
                load8_hilo(a.res.ptr, a.res_lo,
                           (long long)n * a.res.sN + (long long)h * a.res.sH + (long long)w * a.res.sW + c0, r[b]);
        }
        // ---- compute + store -------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < kApplyBatch; ++b) {
            if (state[b] == 0) continue;
            if (state[b] == 1) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[b][q] = 0.f;
                store8_bf16(a.dst, a.dst_lo, doff[b], v[b]);
                continue;
            }
            if (norm) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[b][q] = (v[b][q] - mean[q]) * rstd[q];
            }
            if (a.act == SSCG_ACT_RELU) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[b][q] = fmaxf(v[b][q], 0.f);
            } else if (a.act == SSCG_ACT_LRELU) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[b][q] = v[b][q] > 0.f ? v[b][q] : v[b][q] * a.slope;
            }
            if (seed != 0) {
                const uint32_t bits = drop_bits(seed, (unsigned long long)spix[b] * p.CH + chunk);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[b][q] = ((bits >> q) & 1u) ? 2.f * v[b][q] : 0.f;
            }
            if (has_res) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[b][q] += r[b][q];
            }
            store8_bf16(a.dst, a.dst_lo, doff[b], v[b]);
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

constexpr int kPrepBatch = 2;

__global__ void __launch_bounds__(256, 2) in_bwd_prep_kernel(const __grid_constant__ BwdDev p) {
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
    if (norm && active) load_norm(a.stats, a.eps, (long long)n * a.C + c0, 1.f / (float)(a.H * a.W), mean, rstd);
    const uint64_t seed = (a.drop_seed != 0 && a.drop_ctr) ? (a.drop_seed ^ (*a.drop_ctr * 0x9E3779B97F4A7C15ull))
                                                           : a.drop_seed;
    float acc1[8], acc2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc1[q] = acc2[q] = 0.f;
    const long long npix = (long long)a.H * a.W;
    const long long pix0 = (long long)blockIdx.x * p.rows * p.iters + row;
    if (active) {
        for (int it0 = 0; it0 < p.iters; it0 += kPrepBatch) {
            float g[kPrepBatch][8], z[kPrepBatch][8], sk[kPrepBatch][8];
            long long off[kPrepBatch], spix[kPrepBatch];
            bool live[kPrepBatch];
            // ---- loads -------------------------------------------------------------------------
#pragma unroll
            for (int b = 0; b < kPrepBatch; ++b) {
                const long long pix = pix0 + (long long)(it0 + b) * p.rows;
                live[b] = pix < npix;
                if (!live[b]) continue;
                const int w = pix % a.W, h = pix / a.W;
                spix[b] = (long long)n * npix + pix;
                off[b] = spix[b] * a.C + c0;
                if (a.dyp.ptr != nullptr) {
                    // interior position first (always present), halo positions only near the border
                    load8(a.dyp.ptr, a.dyp_fp32 != 0,
                          (long long)n * a.dyp.sN + (long long)(h + a.pad) * a.dyp.sH + (long long)(w + a.pad) * a.dyp.sW + c0,
                          g[b]);
                    if (fold) {
                        int hq[3], wq[3];
                        const int nh = fold_positions(h, a.H, a.pad, a.pad_mode, hq);
                        const int nw = fold_positions(w, a.W, a.pad, a.pad_mode, wq);
                        if (nh > 1 || nw > 1) {
                            for (int x = 0; x < nh; ++x)
                                for (int y = 0; y < nw; ++y) {
                                    if (x == 0 && y == 0) continue;
                                    float t[8];
                                    load8(a.dyp.ptr, a.dyp_fp32 != 0,
                                          (long long)n * a.dyp.sN + (long long)hq[x] * a.dyp.sH + (long long)wq[y] * a.dyp.sW + c0, t);
#pragma unroll
                                    for (int q = 0; q < 8; ++q) g[b][q] += t[q];
                                }
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[b][q] = 0.f;
                }
                if (a.skip.ptr != nullptr)
                    load8(a.skip.ptr, a.skip_fp32 != 0,
                          (long long)n * a.skip.sN + (long long)h * a.skip.sH + (long long)w * a.skip.sW + c0, sk[b]);
                if (need_raw) load8(a.raw, a.raw_fp32 != 0, off[b], z[b]);
            }
            // ---- compute + store ---------------------------------------------------------------
#pragma unroll
            for (int b = 0; b < kPrepBatch; ++b) {
                if (!live[b]) continue;
                if (a.skip.ptr != nullptr) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[b][q] += sk[b][q];
                }
                if (a.g_out != nullptr) {
                    if (a.g_fp32) store8_f32(a.g_out, off[b], g[b]);
                    else store8_bf16(a.g_out, nullptr, off[b], g[b]);
                }
                if (seed != 0) {
                    const uint32_t bits = drop_bits(seed, (unsigned long long)spix[b] * p.CH + chunk);
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[b][q] = ((bits >> q) & 1u) ? 2.f * g[b][q] : 0.f;
                }
                if (need_raw) {
                    if (norm) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) z[b][q] = (z[b][q] - mean[q]) * rstd[q];
                    }
                    if (a.act == SSCG_ACT_RELU) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[b][q] = z[b][q] > 0.f ? g[b][q] : 0.f;
                    } else if (a.act == SSCG_ACT_LRELU) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[b][q] = z[b][q] > 0.f ? g[b][q] : g[b][q] * a.slope;
                    } else if (a.act == SSCG_ACT_TANH) {   // raw holds y = tanh(.)
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[b][q] = g[b][q] * (1.f - z[b][q] * z[b][q]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) z[b][q] = 0.f;
                }
                if (a.dz_fp32) store8_f32(a.dz, off[b], g[b]);
                else store8_bf16(a.dz, a.dz_lo, off[b], g[b]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    acc1[q] += g[b][q];
                    acc2[q] += g[b][q] * z[b][q];
                }
            }
        }
    }
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
    const int nout = p.CH * 16;    // (chunk, q, {sum, sum*z}) pairs for this block
    for (int o = threadIdx.x; o < nout; o += 256) {
        const int ch = o >> 4, e = o & 15;
        float s = 0.f;
        for (int r = 0; r < p.rows; ++r) s += s_red[(r * p.CH + ch) * 16 + e];
        // e = 2*q + k  ->  bstats[(n*C + ch*8 + q)*2 + k]
        atomicAdd(a.bstats + ((long long)n * a.C + ch * 8) * 2 + e, s);
    }
}

constexpr int kBwdApplyBatch = 4;

__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const __grid_constant__ BwdDev p) {
    const SscgBwdArgs& a = p.a;
    const int n = blockIdx.y;
    const int chunk = threadIdx.x % p.CH;
    const int row = threadIdx.x / p.CH;
    if (row >= p.rows) return;
    const int c0 = chunk * 8;
    const float inv_cnt = 1.f / (float)(a.H * a.W);
    float mean[8], rstd[8], m1[8], m2[8];
    load_norm(a.stats, a.eps, (long long)n * a.C + c0, inv_cnt, mean, rstd);
    {
        const float4* bp = reinterpret_cast<const float4*>(a.bstats + ((long long)n * a.C + c0) * 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 s = bp[q];
            m1[2 * q] = s.x * inv_cnt; m2[2 * q] = s.y * inv_cnt;
            m1[2 * q + 1] = s.z * inv_cnt; m2[2 * q + 1] = s.w * inv_cnt;
        }
    }
    const long long npix = (long long)a.H * a.W;
    const long long pix0 = (long long)blockIdx.x * p.rows * p.iters + row;
    for (int it0 = 0; it0 < p.iters; it0 += kBwdApplyBatch) {
        float z[kBwdApplyBatch][8], g[kBwdApplyBatch][8];
        long long off[kBwdApplyBatch];
        bool live[kBwdApplyBatch];
#pragma unroll
        for (int b = 0; b < kBwdApplyBatch; ++b) {
            const long long pix = pix0 + (long long)(it0 + b) * p.rows;
            live[b] = pix < npix;
            if (!live[b]) continue;
            off[b] = ((long long)n * npix + pix) * a.C + c0;
            load8(a.raw, a.raw_fp32 != 0, off[b], z[b]);
            load8(a.dz, a.dz_fp32 != 0, off[b], g[b]);
        }
#pragma unroll
        for (int b = 0; b < kBwdApplyBatch; ++b) {
            if (!live[b]) continue;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float zz = (z[b][q] - mean[q]) * rstd[q];
                g[b][q] = rstd[q] * (g[b][q] - m1[q] - zz * m2[q]);
            }
            store8_bf16(p.draw, p.draw_lo, off[b], g[b]);
        }
    }
}

