// elementwise.cu — the HBM-bound passes around the tensor-core GEMMs: layout packing at the module
// boundary, InstanceNorm apply (+activation, dropout, residual, halo write), the matching backward
// reductions, and weight-slab preparation.  All of them move 16-byte (8 x bf16) vectors per thread
// with the channel axis innermost, so a warp touches whole 128-byte lines.
#include "sscg_common.cuh"

namespace sscg {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int q, int n) {   // torch ReflectionPad semantics
    if (q < 0) q = -q;
    if (q >= n) q = 2 * (n - 1) - q;
    return q;
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// 8 keep-bits for the 8 channels of one vector; element id = vector index in the unpadded tensor.
// Counter-based: murmur3's 32-bit finaliser over (index, seed) — cheap enough for a bandwidth kernel
// (a 64-bit splitmix cost 20 % of in_apply), identical in forward and backward.
__device__ __forceinline__ uint32_t drop_bits(uint64_t seed, uint64_t vec_index) {
    uint32_t x = static_cast<uint32_t>(vec_index) * 0x9E3779B1u + static_cast<uint32_t>(seed);
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    x ^= x >> 16;
    x ^= static_cast<uint32_t>(seed >> 32) * 0x27D4EB2Fu;
    x ^= x >> 15; x *= 0x2C1B3C6Du;
    x ^= x >> 12;
    return (x >> 11) & 0xffu;
}
__device__ __forceinline__ void load8(const void* base, bool fp32, long long off, float (&v)[8]) {
    if (fp32) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
        const float4 a = p[0], b = p[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[2 * q] = __uint_as_float(w[q] << 16);
            v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
        }
    }
}
// bf16 hi (+ optional lo) planes
__device__ __forceinline__ void load8_hilo(const void* hi, const void* lo, long long off, float (&v)[8]) {
    load8(hi, false, off, v);
    if (lo != nullptr) {
        float l[8];
        load8(lo, false, off, l);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += l[q];
    }
}
__device__ __forceinline__ void store8_bf16(void* hi, void* lo, long long off, const float (&v)[8]) {
    uint32_t h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = pack_bf16x2(v[2 * q], v[2 * q + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(hi) + off) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo != nullptr) {
        uint32_t l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float r0 = v[2 * q] - __uint_as_float(h[q] << 16);
            const float r1 = v[2 * q + 1] - __uint_as_float(h[q] & 0xffff0000u);
            l[q] = pack_bf16x2(r0, r1);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(lo) + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
__device__ __forceinline__ void store8_f32(void* dst, long long off, const float (&v)[8]) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
}
// Plane sums -> per-channel constants of one sample, decoded ONCE per CTA into shared memory (coalesced 32-byte reads,
// one channel per thread and pass) and then picked up by every thread for its 8 channels.  (Per-thread decoding reads
// 256 strided bytes per thread: 8x the L1 wavefronts of the old float statistics, +5 us per launch.)
//   BMEANS = false: accumulators of (sum x, sum x^2)  -> (mean, rstd)
//   BMEANS = true : accumulators of (sum dZ, sum dZ*Z) -> (mean dZ, mean dZ*Z)
// `bar` synchronises the participating threads (all of them call this function); s_pairs holds C float2.
template <bool BMEANS, typename Bar>
__device__ __forceinline__ void cta_load_sums(const void* acc, long long n, int C, float inv_cnt, float eps,
                                              float2* s_pairs, int tid, int nthreads, int c0, bool active, Bar bar,
                                              float (&o1)[8], float (&o2)[8]) {
    bar();                                             // earlier readers are done with s_pairs
    const longlong2* sp = reinterpret_cast<const longlong2*>(reinterpret_cast<const long long*>(acc) + n * C * (2 * kDetWords));
    for (int c = tid; c < C; c += nthreads) {
        const longlong2 s1 = sp[2 * c], s2 = sp[2 * c + 1];
        const float v1 = det_decode(s1.x, s1.y) * inv_cnt;
        float v2 = det_decode(s2.x, s2.y) * inv_cnt;
        if (!BMEANS) v2 = rsqrtf(fmaxf(v2 - v1 * v1, 0.f) + eps);
        s_pairs[c] = make_float2(v1, v2);
    }
    bar();
    if (active) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float2 t = s_pairs[c0 + q];
            o1[q] = t.x;
            o2[q] = t.y;
        }
    }
}
struct BarBlock {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int kThreads>
struct BarNamed1 {
    __device__ __forceinline__ void operator()() const { named_bar_sync(1, kThreads); }
};
constexpr int kMaxNormC = 2048;        // channel bound of the register-batched kernels (checked by the launchers)
constexpr int kMaxStreamC = 512;       // ... of the bulk-pipelined ones (stream_geom: C / 8 <= 64)

// ---------------------------------------------------------------------------------------------
// pack: NCHW fp32 (or int64 labels -> one-hot) -> NHWC bf16 with halo
// ---------------------------------------------------------------------------------------------
template <bool ONEHOT>
__global__ void pack_kernel(const float* __restrict__ src, const long long* __restrict__ labels, int N, int C, int H,
                            int W, void* __restrict__ dst_, void* __restrict__ dst_lo_, int Cp, int pad,
                            int pad_mode, int dst_fp32) {
    pdl_wait();
    pdl_launch();
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(dst_);
    __nv_bfloat16* dst_lo = reinterpret_cast<__nv_bfloat16*>(dst_lo_);
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const long long total = (long long)N * Hp * Wp;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int wp = idx % Wp;
        const int hp = (idx / Wp) % Hp;
        const int n = idx / ((long long)Wp * Hp);
        int h = hp - pad, w = wp - pad;
        bool zero = false;
        if (pad_mode == SSCG_PAD_REFLECT) {
            h = reflect_idx(h, H);
            w = reflect_idx(w, W);
        } else if (h < 0 || h >= H || w < 0 || w >= W) {
            zero = true;
        }
        __nv_bfloat16* d = dst + idx * Cp;
        __nv_bfloat16* dl = (dst_lo && !dst_fp32) ? dst_lo + idx * Cp : nullptr;
        int lab = -1;
        if (ONEHOT && !zero) lab = (int)labels[((long long)n * H + h) * W + w];
        for (int c0 = 0; c0 < Cp; c0 += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c = c0 + q;
                float x = 0.f;
                if (!zero && c < C) {
                    if (ONEHOT) x = (c == lab) ? 1.f : 0.f;
                    else x = src[(((long long)n * C + c) * H + h) * W + w];
                }
                v[q] = x;
            }
            if (dst_fp32) store8_f32(dst_, idx * Cp + c0, v);
            else store8_bf16(d, dl, c0, v);
        }
    }
}

// gradient of the boundary pack: padded NHWC (fp32 or bf16) -> NCHW fp32, folding the halo back
__global__ void unpack_fold_kernel(const void* __restrict__ src, int src_fp32, int N, int C, int H, int W, int Cp,
                                   int pad, int pad_mode, float* __restrict__ dst);

__global__ void unpack_kernel(const float* __restrict__ src, int N, int C, int H, int W, int Cp,
                              float* __restrict__ dst) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)N * H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int w = idx % W;
        const int h = (idx / W) % H;
        const int n = idx / ((long long)W * H);
        const float* s = src + idx * Cp;
        for (int c = 0; c < C; ++c) dst[(((long long)n * C + c) * H + h) * W + w] = s[c];
    }
}

#include "norm_kernels.cuh"
#include "norm_stream.cuh"


__global__ void unpack_fold_kernel(const void* __restrict__ src, int src_fp32, int N, int C, int H, int W, int Cp,
                                   int pad, int pad_mode, float* __restrict__ dst) {
    pdl_wait();
    pdl_launch();
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const long long total = (long long)N * H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int w = idx % W;
        const int h = (idx / W) % H;
        const int n = idx / ((long long)W * H);
        int hq[3], wq[3];
        const int nh = fold_positions(h, H, pad, pad_mode, hq);
        const int nw = fold_positions(w, W, pad, pad_mode, wq);
        for (int c0 = 0; c0 < C; c0 += 8) {
            float g[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) g[q] = 0.f;
            for (int x = 0; x < nh; ++x)
                for (int y = 0; y < nw; ++y) {
                    float t[8];
                    load8(src, src_fp32 != 0, (((long long)n * Hp + hq[x]) * Wp + wq[y]) * Cp + c0, t);
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] += t[q];
                }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (c0 + q < C) dst[(((long long)n * C + c0 + q) * H + h) * W + w] = g[q];
        }
    }
}

// bias gradient of a conv without normalisation: grad[c] += scale * sum_n bstats[n][c][0]
__global__ void bias_grad_kernel(const long long* __restrict__ bstats, int N, int C, int Cp, float* __restrict__ grad,
                                 float scale) {
    pdl_wait();
    pdl_launch();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int n = 0; n < N; ++n) {
        const long long* b = bstats + ((long long)n * Cp + c) * (2 * kDetWords);
        s += det_decode(b[0], b[1]);
    }
    grad[c] += scale * s;
}

// ---------------------------------------------------------------------------------------------
// weight slabs
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long wslab_src_index(const SscgWprepArgs& a, int t, int r, int k) {
    int co, ci, kh, kw;
    if (a.mode == 3 || a.mode == 4) {
        // N-expanded 7x7 slabs (conv_nexp.cu): t = kh * n_ntiles + nt, row r = kw * CoW + local column (CoW = a.Cp)
        const int cow = a.Cp;
        const int nn = ((a.mode == 3 ? a.Co : a.Ci) + cow - 1) / cow;
        const int nt = t % nn;
        kh = t / nn;
        kw = r / cow;
        const int loc = nt * cow + (r - kw * cow);
        if (kw >= a.KW) return -1;
        if (a.mode == 3) { co = loc; ci = k; }
        else { ci = loc; co = k; kh = a.KH - 1 - kh; kw = a.KW - 1 - kw; }     // data gradient: flipped taps
    } else if (a.mode == 1) {
        kh = t; kw = k / a.Cp; ci = k - kw * a.Cp; co = r;
        if (kw >= a.KW) return -1;
    } else if (a.mode == 5) {
        // pixel-row order: K = 64 * g + 8 * kw + c8 (channel group g of 8 channels, 8 pixels per group row)
        const int g = k >> 6, rr = k & 63;
        kh = t; kw = rr >> 3; ci = g * 8 + (rr & 7); co = r;
        if (kw >= a.KW) return -1;
    } else {
        kh = t / a.KW; kw = t - kh * a.KW;
        if (a.mode == 0) { co = r; ci = k; } else { ci = r; co = k; }
    }
    if (co >= a.Co || ci >= a.Ci) return -1;
    return a.transposed ? ((((long long)ci * a.Co + co) * a.KH + kh) * a.KW + kw)
                        : ((((long long)co * a.Ci + ci) * a.KH + kh) * a.KW + kw);
}
__host__ __device__ __forceinline__ int wslab_ntaps(const SscgWprepArgs& a) {
    if (a.mode == 3) return a.KH * ((a.Co + a.Cp - 1) / a.Cp);
    if (a.mode == 4) return a.KH * ((a.Ci + a.Cp - 1) / a.Cp);
    return (a.mode == 1 || a.mode == 5) ? a.KH : a.KH * a.KW;
}

__global__ void wprep_kernel(const __grid_constant__ SscgWprepArgs a) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)wslab_ntaps(a) * a.rows_pad * a.Kc;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = idx % a.Kc;
        const int r = (idx / a.Kc) % a.rows_pad;
        const int t = idx / ((long long)a.Kc * a.rows_pad);
        const long long s = wslab_src_index(a, t, r, k);
        const float v = s >= 0 ? a.w[s] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        reinterpret_cast<__nv_bfloat16*>(a.dst)[idx] = h;
        if (a.dst_lo != nullptr)
            reinterpret_cast<__nv_bfloat16*>(a.dst_lo)[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

__global__ void wgrad_unpack_kernel(const __grid_constant__ SscgWprepArgs a, const float* __restrict__ slab,
                                    float* __restrict__ grad, float scale) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)wslab_ntaps(a) * a.rows_pad * a.Kc;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = idx % a.Kc;
        const int r = (idx / a.Kc) % a.rows_pad;
        const int t = idx / ((long long)a.Kc * a.rows_pad);
        const long long s = wslab_src_index(a, t, r, k);
        if (s >= 0) grad[s] += scale * slab[idx];
    }
}

// Batched variants: one launch walks a device table of slab descriptors (all stages of a network), instead
// of one ~10 us launch per stage and slab (116 weight preparations + 58 gradient re-layouts per training step).
__device__ __forceinline__ int wbatch_find(const SscgWbatchEntry* __restrict__ tab, int count, long long idx) {
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].start <= idx) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// A thread owns one (row, k) position of a slab and walks the taps: for the fixed (co, ci) pair the taps are
// adjacent in the PyTorch weight layout, so the strided 4-byte gathers of a warp hit the same sectors on every
// tap (L1) instead of touching 9-49x the useful bytes, and every store of a warp is one contiguous run.
// `start` / `total` count (row, k) positions: rows_pad * Kc per entry.
__global__ void wprep_batch_kernel(const SscgWbatchEntry* __restrict__ tab, int count, long long total) {
    pdl_wait();
    pdl_launch();
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < total;
         gidx += (long long)gridDim.x * blockDim.x) {
        const SscgWbatchEntry& e = tab[wbatch_find(tab, count, gidx)];
        const SscgWprepArgs& a = e.a;
        const long long idx = gidx - e.start;
        const int k = idx % a.Kc;
        const int r = idx / a.Kc;
        const int nt = wslab_ntaps(a);
        const long long plane = (long long)a.rows_pad * a.Kc;
        for (int t = 0; t < nt; ++t) {
            const long long s = wslab_src_index(a, t, r, k);
            const float v = s >= 0 ? a.w[s] : 0.f;
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            reinterpret_cast<__nv_bfloat16*>(a.dst)[t * plane + idx] = h;
            if (a.dst_lo != nullptr)
                reinterpret_cast<__nv_bfloat16*>(a.dst_lo)[t * plane + idx] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}
__global__ void wgrad_unpack_batch_kernel(const SscgWbatchEntry* __restrict__ tab, int count, long long total, float scale) {
    pdl_wait();
    pdl_launch();
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < total;
         gidx += (long long)gridDim.x * blockDim.x) {
        const SscgWbatchEntry& e = tab[wbatch_find(tab, count, gidx)];
        const SscgWprepArgs& a = e.a;
        const long long idx = gidx - e.start;
        const int k = idx % a.Kc;
        const int r = idx / a.Kc;
        const int nt = wslab_ntaps(a);
        const long long plane = (long long)a.rows_pad * a.Kc;
        for (int t = 0; t < nt; ++t) {
            const long long s = wslab_src_index(a, t, r, k);
            if (s >= 0) e.grad[s] += scale * e.slab[t * plane + idx];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bilinear resize, align_corners = True, NCHW fp32 — nn.Upsample(size, mode='bilinear', align_corners=True), the
// reference's `interp` (model.py:62-63,268; applied at model.py:132,390-392,413-415,562,581-594).  Same index
// arithmetic and operation order as ATen's upsample_bilinear2d (source index = dst * (in - 1) / (out - 1) in fp32).
// The backward is a GATHER over the output gradient (each input element sums the outputs that read it, in raster
// order): no atomics, reproducible.
// ---------------------------------------------------------------------------------------------
struct InterpTap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ InterpTap interp_tap(int o, float scale, int in_size) {
    InterpTap t;
    const float src = scale * (float)o;
    t.i0 = (int)src;                                   // src >= 0
    if (t.i0 > in_size - 1) t.i0 = in_size - 1;
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = src - (float)t.i0;
    t.l0 = 1.f - t.l1;
    return t;
}
__global__ void __launch_bounds__(256) interp_fwd_kernel(const float* __restrict__ x, int NC, int Hi, int Wi,
                                                         float* __restrict__ y, int Ho, int Wo, float sh, float sw) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)NC * Ho * Wo;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int ow = idx % Wo;
        const int oh = (idx / Wo) % Ho;
        const long long nc = idx / ((long long)Wo * Ho);
        const InterpTap th = interp_tap(oh, sh, Hi), tw = interp_tap(ow, sw, Wi);
        const float* p = x + nc * Hi * Wi;
        y[idx] = th.l0 * (tw.l0 * p[th.i0 * Wi + tw.i0] + tw.l1 * p[th.i0 * Wi + tw.i1]) +
                 th.l1 * (tw.l0 * p[th.i1 * Wi + tw.i0] + tw.l1 * p[th.i1 * Wi + tw.i1]);
    }
}
__global__ void __launch_bounds__(256) interp_bwd_kernel(const float* __restrict__ dy, int NC, int Hi, int Wi,
                                                         float* __restrict__ dx, int Ho, int Wo, float sh, float sw) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)NC * Hi * Wi;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int iw = idx % Wi;
        const int ih = (idx / Wi) % Hi;
        const long long nc = idx / ((long long)Wi * Hi);
        // candidate outputs: those whose source index lies within one input pixel of (ih, iw); tested exactly below
        int oh0 = 0, oh1 = Ho - 1, ow0 = 0, ow1 = Wo - 1;
        if (sh > 0.f) {
            oh0 = max(0, (int)floorf((float)(ih - 1) / sh) - 1);
            oh1 = min(Ho - 1, (int)ceilf((float)(ih + 1) / sh) + 1);
        }
        if (sw > 0.f) {
            ow0 = max(0, (int)floorf((float)(iw - 1) / sw) - 1);
            ow1 = min(Wo - 1, (int)ceilf((float)(iw + 1) / sw) + 1);
        }
        const float* g = dy + nc * Ho * Wo;
        float acc = 0.f;
        for (int oh = oh0; oh <= oh1; ++oh) {
            const InterpTap th = interp_tap(oh, sh, Hi);
            float wh = 0.f;
            if (th.i0 == ih) wh += th.l0;
            if (th.i1 == ih) wh += th.l1;            // i1 == i0 at the last row: both weights land on it (l1 = 0 there)
            if (wh == 0.f && th.i0 != ih && th.i1 != ih) continue;
            for (int ow = ow0; ow <= ow1; ++ow) {
                const InterpTap tw = interp_tap(ow, sw, Wi);
                float ww = 0.f;
                if (tw.i0 == iw) ww += tw.l0;
                if (tw.i1 == iw) ww += tw.l1;
                if (tw.i0 != iw && tw.i1 != iw) continue;
                acc += wh * ww * g[oh * Wo + ow];
            }
        }
        dx[idx] = acc;
    }
}

static inline int ew_grid(long long total, int block) {
    long long g = (total + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return (int)g;
}

// rows/iters choice for the (C/8)-vector x pixel-row block layout
static void vec_layout(int C, long long npix, int& CH, int& rows, int& iters, int& gridx) {
    CH = C / 8;
    rows = 256 / CH;
    if (rows < 1) rows = 1;
    // pixels per thread: enough to amortise the per-thread statistics setup and the per-block atomics,
    // while keeping >= ~2 blocks per SM in flight for a 16-sample batch (grid.y = N multiplies this)
    iters = 32;
    while (iters > 8 && npix / ((long long)rows * iters) < 16) iters /= 2;
    long long per_block = (long long)rows * iters;
    gridx = (int)((npix + per_block - 1) / per_block);
    if (gridx < 1) gridx = 1;
}

// ---- bulk-pipelined variants (norm_stream.cuh): geometry + eligibility -----------------------------
static int g_stream_norm = -1;     // -1: read SSCG_STREAM_NORM from the environment on first use
static bool stream_enabled() {
    if (g_stream_norm < 0) {
        const char* e = getenv("SSCG_STREAM_NORM");
        g_stream_norm = (e && e[0] >= '0' && e[0] <= '2') ? (e[0] - '0') : 2;
    }
    return g_stream_norm != 0;
}
// Mode 1 keeps the register-batched first backward half (the pipelined one is 8-20 % faster at production
// shapes: 39 vs 43 us on 16 x 64 x 64 x 256 with a reflect halo); mode 2, the default, pipelines all three.
static bool stream_prep_enabled() { return stream_enabled() && g_stream_norm >= 2; }
static int ew_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
// Unit = 1/d of an image row (d = smallest divisor of W that brings the unit under kStrUnitMax; a whole
// row of up to 48 KB when W has no such divisor).  Returns false when the shape does not suit the ring.
static bool stream_geom(int N, int H, int W, int C, int ntens, int threads, StreamGeom& g, int budget = kStrSmemBudget,
                        int halo_px = 0) {
    if (!stream_enabled() || C % 8 || N < 1 || H < 1 || W < 1) return false;
    const int CH = C / 8;
    if (CH > 64 || threads % CH) return false;
    // the stage descriptors carry 32-bit byte offsets into the (haloed) tensors
    if ((long long)N * (H + 16) * (W + 16) * C * 2 >= (1LL << 32)) return false;
    const long long row = (long long)W * C * 2;
    long long unit_max = budget / (3 * ntens);       // at least three stages in the ring
    if (unit_max > kStrUnitMax) unit_max = kStrUnitMax;
    int d = 0;
    for (int t = 1; t <= W; ++t)
        if (W % t == 0 && row / t <= unit_max) { d = t; break; }
    if (d == 0 || row / d < 4096) {
        if (row > 48 * 1024 || row < 4096) return false;
        d = 1;
    }
    g.upr = d;
    g.seg_px = W / d;
    g.ub = g.seg_px * C * 2;
    g.ups = H * d;
    g.total = N * g.ups;
    g.ntens = ntens;
    g.CH = CH;
    g.slot0 = ((g.seg_px + 2 * halo_px) * C * 2 + 127) & ~127;
    g.stage_bytes = g.slot0 + (ntens - 1) * g.ub;
    int nst = budget / g.stage_bytes;
    if (nst > kStrMaxStages) nst = kStrMaxStages;
    if (nst < 2) return false;
    g.nst = nst;
    return true;
}
template <typename KernelT>
static int stream_prepare(KernelT kernel, int& done) {
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStrSmemBudget + 1024);
    if (e != cudaSuccess) return set_error("norm stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    done = 1;
    return 0;
}
static inline size_t stream_smem(const StreamGeom& g) { return (size_t)g.nst * g.stage_bytes + 1024; }   // alignment + barriers + stage descriptors
static inline int stream_grid(const StreamGeom& g) { return g.total < ew_sm_count() ? g.total : ew_sm_count(); }

}  // namespace sscg

using namespace sscg;

#define SSCG_CHECK_LAUNCH(name)                                                             \
    do {                                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                \
        if (e_ != cudaSuccess) return set_error(name " launch: %s", cudaGetErrorString(e_)); \
    } while (0)

extern "C" int sscg_pack_nchw(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, void* dst, void* dst_lo,
                              int32_t dst_fp32, int32_t Cp, int32_t pad, int32_t pad_mode, void* stream) {
    if (Cp % 8 || Cp < C) return set_error("pack_nchw: Cp=%d must be a multiple of 8 and >= C=%d", Cp, C);
    const long long total = (long long)N * (H + 2 * pad) * (W + 2 * pad);
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(pack_kernel<false>, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
            src, nullptr, N, C, H, W, dst, dst_lo, Cp, pad, pad_mode, dst_fp32);
    }
    SSCG_CHECK_LAUNCH("pack_nchw");
    return 0;
}

extern "C" int sscg_onehot_pack(const int64_t* labels, int32_t N, int32_t C, int32_t H, int32_t W, void* dst,
                                void* dst_lo, int32_t Cp, int32_t pad, int32_t pad_mode, void* stream) {
    if (Cp % 8 || Cp < C) return set_error("onehot_pack: Cp=%d must be a multiple of 8 and >= C=%d", Cp, C);
    const long long total = (long long)N * (H + 2 * pad) * (W + 2 * pad);
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(pack_kernel<true>, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
            nullptr, reinterpret_cast<const long long*>(labels), N, C, H, W, dst, dst_lo, Cp, pad, pad_mode, 0);
    }
    SSCG_CHECK_LAUNCH("onehot_pack");
    return 0;
}

extern "C" int sscg_unpack_nhwc(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cp, float* dst,
                                void* stream) {
    const long long total = (long long)N * H * W;
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(unpack_kernel, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), src, N, C, H, W, Cp, dst);
    }
    SSCG_CHECK_LAUNCH("unpack_nhwc");
    return 0;
}

extern "C" int sscg_unpack_fold(const void* src, int32_t src_fp32, int32_t N, int32_t C, int32_t H, int32_t W,
                                int32_t Cp, int32_t pad, int32_t pad_mode, float* dst, void* stream) {
    if (Cp % 8) return set_error("unpack_fold: Cp=%d must be a multiple of 8", Cp);
    const long long total = (long long)N * H * W;
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(unpack_fold_kernel, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), src, src_fp32, N, C, H, W, Cp,
                                                                                             pad, pad_mode, dst);
    }
    SSCG_CHECK_LAUNCH("unpack_fold");
    return 0;
}

extern "C" int sscg_bias_grad(const void* bstats, int32_t N, int32_t C, int32_t Cp, float* grad, float scale,
                              void* stream) {
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(bias_grad_kernel, (C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), 
            reinterpret_cast<const long long*>(bstats), N, C, Cp, grad, scale);
    }
    SSCG_CHECK_LAUNCH("bias_grad");
    return 0;
}

extern "C" int sscg_set_stream_norm(int32_t on) {
    g_stream_norm = on < 0 ? 0 : (on > 2 ? 2 : on);
    return 0;
}

extern "C" int sscg_in_apply(const SscgApplyArgs* a, void* stream) {
    if (a->C % 8 || a->C > 2048) return set_error("in_apply: C=%d must be a multiple of 8 (<= 2048)", a->C);
    {
        // bulk-pipelined fast path: all-bf16, halo (if any) written by reflection
        ApplyStreamDev sd;
        const bool plain = !(a->raw_fp32 || a->res_lo || a->dst_lo);
        const bool halo_ok = a->pad == 0 || a->pad_mode == SSCG_PAD_REFLECT;
        const bool res_ok = a->res.ptr == nullptr || a->res.sW == a->C;
        if (plain && halo_ok && res_ok && a->pad < a->H && a->pad < a->W &&
            stream_geom(a->N, a->H, a->W, a->C, a->res.ptr ? 2 : 1, kStrThreadsLight, sd.g)) {
            static int prepared[4] = {0, 0, 0, 0};
            if (int rc = stream_prepare(in_apply_stream_kernel<kStrThreadsLight, 0>, prepared[0])) return rc;
            if (int rc = stream_prepare(in_apply_stream_kernel<kStrThreadsLight, 1>, prepared[1])) return rc;
            if (int rc = stream_prepare(in_apply_stream_kernel<kStrThreadsLight, 2>, prepared[2])) return rc;
            if (int rc = stream_prepare(in_apply_stream_kernel<kStrThreadsLight, 3>, prepared[3])) return rc;
            sd.a = *a;
            int spec = 0;
            if (a->stats) {
                if (a->act == SSCG_ACT_RELU && !a->res.ptr) spec = a->drop_seed != 0 ? 2 : 1;
                else if (a->act == SSCG_ACT_NONE && a->res.ptr && a->drop_seed == 0) spec = 3;
            }
            {
                LaunchScope ls_(7, static_cast<cudaStream_t>(stream));
                const dim3 grid(stream_grid(sd.g)), block(kStrThreadsLight + 32);
                const size_t smem = stream_smem(sd.g);
                cudaStream_t st = static_cast<cudaStream_t>(stream);
                if (spec == 1) launch_k(in_apply_stream_kernel<kStrThreadsLight, 1>, grid, block, smem, st, sd);
                else if (spec == 2) launch_k(in_apply_stream_kernel<kStrThreadsLight, 2>, grid, block, smem, st, sd);
                else if (spec == 3) launch_k(in_apply_stream_kernel<kStrThreadsLight, 3>, grid, block, smem, st, sd);
                else launch_k(in_apply_stream_kernel<kStrThreadsLight, 0>, grid, block, smem, st, sd);
            }
            SSCG_CHECK_LAUNCH("in_apply_stream");
            return 0;
        }
    }
    ApplyDev d;
    d.a = *a;
    int gridx;
    vec_layout(a->C, (long long)(a->H + 2 * a->pad) * (a->W + 2 * a->pad), d.CH, d.rows, d.iters, gridx);
    {
        LaunchScope ls_(7, static_cast<cudaStream_t>(stream));
        if (a->raw_fp32 || a->res_lo || a->dst_lo)
            launch_k(in_apply_kernel<true>, dim3(gridx, a->N), 256, 0, static_cast<cudaStream_t>(stream), d);
        else
            launch_k(in_apply_kernel<false>, dim3(gridx, a->N), 256, 0, static_cast<cudaStream_t>(stream), d);
    }
    SSCG_CHECK_LAUNCH("in_apply");
    return 0;
}

extern "C" int sscg_in_bwd_prep(const SscgBwdArgs* a, void* stream) {
    if (a->C % 8 || a->C > 2048) return set_error("in_bwd_prep: C=%d must be a multiple of 8 (<= 2048)", a->C);
    {
        BwdStreamDev sd;
        const bool plain = !(a->raw_fp32 || a->dyp_fp32 || a->skip_fp32 || a->g_fp32 || a->dz_fp32 || a->dz_lo);
        const bool need_raw = a->stats != nullptr || a->act != SSCG_ACT_NONE;
        const bool views_ok = a->dyp.ptr != nullptr && a->dyp.sW == a->C && (a->skip.ptr == nullptr || a->skip.sW == a->C);
        const bool pad_ok = a->pad == 0 || (a->pad < a->H && a->pad < a->W);
        const int ntens = 1 + (a->skip.ptr ? 1 : 0) + (need_raw ? 1 : 0);
        if (plain && views_ok && pad_ok && a->dz != nullptr && stream_prep_enabled() && stream_geom(a->N, a->H, a->W, a->C, ntens, kStrThreadsHeavy, sd.g, kStrSmemBudget - kStrThreadsHeavy * 16 * 4,
                        (a->pad_mode == SSCG_PAD_REFLECT) ? a->pad : 0)) {
            static int prepared[3] = {0, 0, 0};
            if (int rc = stream_prepare(in_bwd_prep_stream_kernel<kStrThreadsHeavy, 0>, prepared[0])) return rc;
            if (int rc = stream_prepare(in_bwd_prep_stream_kernel<kStrThreadsHeavy, 1>, prepared[1])) return rc;
            if (int rc = stream_prepare(in_bwd_prep_stream_kernel<kStrThreadsHeavy, 2>, prepared[2])) return rc;
            sd.a = *a; sd.draw = nullptr;
            int spec = 0;
            if (a->stats && a->bstats) {
                if (a->act == SSCG_ACT_RELU && !a->skip.ptr && !a->g_out) spec = 1;
                else if (a->act == SSCG_ACT_NONE && a->skip.ptr && a->g_out && a->drop_seed == 0) spec = 2;
            }
            {
                LaunchScope ls_(8, static_cast<cudaStream_t>(stream));
                const dim3 grid(stream_grid(sd.g)), block(kStrThreadsHeavy + 32);
                const size_t smem = stream_smem(sd.g);
                cudaStream_t st = static_cast<cudaStream_t>(stream);
                if (spec == 1) launch_k(in_bwd_prep_stream_kernel<kStrThreadsHeavy, 1>, grid, block, smem, st, sd);
                else if (spec == 2) launch_k(in_bwd_prep_stream_kernel<kStrThreadsHeavy, 2>, grid, block, smem, st, sd);
                else launch_k(in_bwd_prep_stream_kernel<kStrThreadsHeavy, 0>, grid, block, smem, st, sd);
            }
            SSCG_CHECK_LAUNCH("in_bwd_prep_stream");
            return 0;
        }
    }
    BwdDev d;
    d.a = *a; d.draw = nullptr; d.draw_lo = nullptr;
    int gridx;
    vec_layout(a->C, (long long)a->H * a->W, d.CH, d.rows, d.iters, gridx);
    {
        LaunchScope ls_(8, static_cast<cudaStream_t>(stream));
        if (a->raw_fp32 || a->dyp_fp32 || a->skip_fp32 || a->g_fp32 || a->dz_fp32 || a->dz_lo)
            launch_k(in_bwd_prep_kernel<true>, dim3(gridx, a->N), 256, 0, static_cast<cudaStream_t>(stream), d);
        else
            launch_k(in_bwd_prep_kernel<false>, dim3(gridx, a->N), 256, 0, static_cast<cudaStream_t>(stream), d);
    }
    SSCG_CHECK_LAUNCH("in_bwd_prep");
    return 0;
}

extern "C" int sscg_in_bwd_apply(const SscgBwdArgs* a, void* draw, void* draw_lo, void* stream) {
    if (a->C % 8 || a->C > 2048) return set_error("in_bwd_apply: C=%d must be a multiple of 8 (<= 2048)", a->C);
    if (!a->stats || !a->bstats) return set_error("in_bwd_apply: needs stats and bstats");
    {
        BwdStreamDev sd;
        if (!(a->raw_fp32 || a->dz_fp32 || draw_lo) && stream_geom(a->N, a->H, a->W, a->C, 2, kStrThreadsLight, sd.g)) {
            static int prepared = 0;
            if (int rc = stream_prepare(in_bwd_apply_stream_kernel<kStrThreadsLight>, prepared)) return rc;
            sd.a = *a; sd.draw = draw;
            {
                LaunchScope ls_(8, static_cast<cudaStream_t>(stream));
                launch_k(in_bwd_apply_stream_kernel<kStrThreadsLight>, stream_grid(sd.g), kStrThreadsLight + 32, stream_smem(sd.g),
                                             static_cast<cudaStream_t>(stream), sd);
            }
            SSCG_CHECK_LAUNCH("in_bwd_apply_stream");
            return 0;
        }
    }
    BwdDev d;
    d.a = *a; d.draw = draw; d.draw_lo = draw_lo;
    int gridx;
    vec_layout(a->C, (long long)a->H * a->W, d.CH, d.rows, d.iters, gridx);
    {
        LaunchScope ls_(8, static_cast<cudaStream_t>(stream));
        if (a->raw_fp32 || a->dz_fp32 || draw_lo)
            launch_k(in_bwd_apply_kernel<true>, dim3(gridx, a->N), 256, 0, static_cast<cudaStream_t>(stream), d);
        else
            launch_k(in_bwd_apply_kernel<false>, dim3(gridx, a->N), 256, 0, static_cast<cudaStream_t>(stream), d);
    }
    SSCG_CHECK_LAUNCH("in_bwd_apply");
    return 0;
}

static inline float interp_scale(int in_size, int out_size) {
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}
extern "C" int sscg_interp_bilinear_fwd(const float* x, int32_t N, int32_t C, int32_t Hi, int32_t Wi, float* y, int32_t Ho,
                                        int32_t Wo, void* stream) {
    if (!x || !y || N < 1 || C < 1 || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1) return set_error("interp_bilinear_fwd: bad arguments");
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(interp_fwd_kernel, ew_grid((long long)N * C * Ho * Wo, 256), 256, 0, static_cast<cudaStream_t>(stream), 
            x, N * C, Hi, Wi, y, Ho, Wo, interp_scale(Hi, Ho), interp_scale(Wi, Wo));
    }
    SSCG_CHECK_LAUNCH("interp_bilinear_fwd");
    return 0;
}
extern "C" int sscg_interp_bilinear_bwd(const float* dy, int32_t N, int32_t C, int32_t Hi, int32_t Wi, int32_t Ho,
                                        int32_t Wo, float* dx, void* stream) {
    if (!dy || !dx || N < 1 || C < 1 || Hi < 1 || Wi < 1 || Ho < 1 || Wo < 1) return set_error("interp_bilinear_bwd: bad arguments");
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(interp_bwd_kernel, ew_grid((long long)N * C * Hi * Wi, 256), 256, 0, static_cast<cudaStream_t>(stream), 
            dy, N * C, Hi, Wi, dx, Ho, Wo, interp_scale(Hi, Ho), interp_scale(Wi, Wo));
    }
    SSCG_CHECK_LAUNCH("interp_bilinear_bwd");
    return 0;
}

extern "C" int sscg_wprep(const SscgWprepArgs* a, void* stream) {
    const long long total = (long long)wslab_ntaps(*a) * a->rows_pad * a->Kc;
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(wprep_kernel, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), *a);
    }
    SSCG_CHECK_LAUNCH("wprep");
    return 0;
}

extern "C" int sscg_wgrad_unpack(const SscgWprepArgs* a, const float* slab, float* grad, float scale, void* stream) {
    const long long total = (long long)wslab_ntaps(*a) * a->rows_pad * a->Kc;
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(wgrad_unpack_kernel, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), *a, slab, grad, scale);
    }
    SSCG_CHECK_LAUNCH("wgrad_unpack");
    return 0;
}

extern "C" int sscg_wprep_batch(const SscgWbatchEntry* table_dev, int32_t count, int64_t total, void* stream) {
    if (!table_dev || count < 1 || total < 1) return set_error("wprep_batch: bad arguments");
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(wprep_batch_kernel, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), table_dev, count, total);
    }
    SSCG_CHECK_LAUNCH("wprep_batch");
    return 0;
}

extern "C" int sscg_wgrad_unpack_batch(const SscgWbatchEntry* table_dev, int32_t count, int64_t total, float scale,
                                       void* stream) {
    if (!table_dev || count < 1 || total < 1) return set_error("wgrad_unpack_batch: bad arguments");
    {
        LaunchScope ls_(9, static_cast<cudaStream_t>(stream));
        launch_k(wgrad_unpack_batch_kernel, ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream), table_dev, count, total,
                                                                                                  scale);
    }
    SSCG_CHECK_LAUNCH("wgrad_unpack_batch");
    return 0;
}
