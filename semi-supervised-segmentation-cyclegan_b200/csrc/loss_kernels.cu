// loss_kernels.cu — fused segmentation-head loss: softmax over the class axis, cross-entropy against
// the label map and the first-max argmax in ONE pass over the NCHW logits, and the combined backward
// (cross-entropy gradient + softmax Jacobian of an incoming probability gradient) in one pass.
//
// Replaces (reference): nn.CrossEntropyLoss = log_softmax + nll_loss2d (model.py:272,398,455),
// nn.Softmax2d (model.py:273,401-402) and `.max(1)[1]` (model.py:435,509) — six ATen kernels forward
// and four backward per logits tensor.  One thread owns one pixel and walks the C class planes, so
// every global access of a warp is a contiguous 128-byte row of one class plane.
#include "sscg_common.cuh"

namespace sscg {

constexpr int kMaxClasses = 32;

// ---------------------------------------------------------------------------------------------
// Grid-wide sums without floating-point atomics: every block stores its (up to two) partial sums to
// its own slot of the workspace, the block that arrives last (sscg_ptx.cuh) adds the slots in a fixed
// pattern — thread t takes slots t, t + 256, ... in order, then a fixed shuffle / shared-memory tree —
// and writes the results.  ws: [0] arrival counter (zero between launches), [4..) float2 slots.
// ---------------------------------------------------------------------------------------------
constexpr int kLossWsHeader = 4;        // floats reserved in front of the slots (counter + padding)
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* s_w) {
    for (int off = 16; off >= 1; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = make_float2(a, b);
    __syncthreads();
    float2 t = make_float2(0.f, 0.f);
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { t.x += s_w[i].x; t.y += s_w[i].y; }
    return t;                            // valid in thread 0
}
// out[0] = scale * sum of a over the grid, out[1] = sum of b (when out1 is set)
__device__ __forceinline__ void grid_sum2_store(float a, float b, float* ws, float* out, float scale, bool out1) {
    __shared__ float2 s_w[8];
    __shared__ unsigned int s_last;
    const float2 t = block_sum2(a, b, s_w);
    float2* slots = reinterpret_cast<float2*>(ws + kLossWsHeader);
    if (threadIdx.x == 0) {
        slots[blockIdx.x] = t;
        unsigned int* ctr = reinterpret_cast<unsigned int*>(ws);
        const unsigned int old = atom_add_acq_rel_gpu(ctr, 1u);
        const bool last = (old + 1u == gridDim.x);
        if (last) {
            *ctr = 0u;
            __threadfence();
        }
        s_last = last ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    float sa = 0.f, sb = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        const float2 v = ld_cg_f2(reinterpret_cast<const float*>(slots + i));
        sa += v.x;
        sb += v.y;
    }
    __syncthreads();                     // s_w is reused
    const float2 r = block_sum2(sa, sb, s_w);
    if (threadIdx.x == 0) {
        out[0] = r.x * scale;
        if (out1) out[1] = r.y;
    }
}

// logits [N][C][H][W] fp32; labels [N][H][W] int64 (or null); probs [N][C][H][W] (or null);
// argmax [N][H][W] int64 (or null); loss_out[2] = (sum of -log p[label] over the counted pixels, their number).
// nn.CrossEntropyLoss semantics (model.py:272): pixels whose label equals ignore_index are not counted; any other label
// outside [0, C) raises the device error flag (torch device-asserts there) and is not counted either.
// V consecutive pixels per thread (vector loads of 4 * V bytes per class plane: one warp instruction then covers
// 128 * V contiguous bytes of a plane instead of 128 — the one-pixel version ran at 1.5-2.2 TB/s); the arithmetic per
// pixel is unchanged.  V > 1 needs HW % V == 0 and 4 * V-byte aligned planes (checked by the launchers).
template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&o)[V]);
template <>
__device__ __forceinline__ void ldv<1>(const float* p, float (&o)[1]) { o[0] = *p; }
template <>
__device__ __forceinline__ void ldv<2>(const float* p, float (&o)[2]) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    o[0] = t.x; o[1] = t.y;
}
template <>
__device__ __forceinline__ void ldv<4>(const float* p, float (&o)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float (&o)[V]);
template <>
__device__ __forceinline__ void stv<1>(float* p, const float (&o)[1]) { *p = o[0]; }
template <>
__device__ __forceinline__ void stv<2>(float* p, const float (&o)[2]) { *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]); }
template <>
__device__ __forceinline__ void stv<4>(float* p, const float (&o)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}

template <int V>
__global__ void __launch_bounds__(256) seg_head_fwd_kernel(const float* __restrict__ logits,
                                                           const long long* __restrict__ labels, int N, int C,
                                                           long long HW, long long ignore_index,
                                                           float* __restrict__ probs, long long* __restrict__ argmax,
                                                           float* __restrict__ loss_out, float* __restrict__ ws) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)N * HW / V;          // groups of V pixels (a group never straddles samples)
    float local = 0.f, counted = 0.f;
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < total;
         gidx += (long long)gridDim.x * blockDim.x) {
        const long long idx = gidx * V;
        const long long n = idx / HW, pix = idx - n * HW;
        const float* src = logits + n * C * HW + pix;
        // three streaming passes over the class planes (max / argmax, sum of exponentials, probabilities); the second and
        // third hit the caches.  Holding the 21 logits of the pixels in registers limits the resident threads instead.
        float mx[V];
        int am[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { mx[i] = -INFINITY; am[i] = 0; }
#pragma unroll 7
        for (int c = 0; c < C; ++c) {
            float v[V];
            ldv<V>(src + c * HW, v);
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (v[i] > mx[i]) { mx[i] = v[i]; am[i] = c; }      // strict '>' keeps the FIRST maximum (torch.max rule)
        }
        float vl[V];
        bool count_it[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            vl[i] = 0.f;
            count_it[i] = false;
            if (labels != nullptr && loss_out != nullptr) {
                const long long lab = labels[idx + i];
                if (lab >= 0 && lab < C) {
                    count_it[i] = true;
                    vl[i] = src[lab * HW + i];
                } else if (lab != ignore_index) {
                    atomicCAS(&g_sscg_dev_error, 0u, (31u << 16) | 0x80000000u);     // label outside [0, C)
                }
            }
        }
        float sum[V];
#pragma unroll
        for (int i = 0; i < V; ++i) sum[i] = 0.f;
#pragma unroll 7
        for (int c = 0; c < C; ++c) {
            float v[V];
            ldv<V>(src + c * HW, v);
#pragma unroll
            for (int i = 0; i < V; ++i) sum[i] += __expf(v[i] - mx[i]);
        }
        if (probs != nullptr) {
            float inv[V];
#pragma unroll
            for (int i = 0; i < V; ++i) inv[i] = 1.f / sum[i];
            float* dst = probs + n * C * HW + pix;
#pragma unroll 7
            for (int c = 0; c < C; ++c) {
                float v[V], o[V];
                ldv<V>(src + c * HW, v);
#pragma unroll
                for (int i = 0; i < V; ++i) o[i] = __expf(v[i] - mx[i]) * inv[i];
                stv<V>(dst + c * HW, o);
            }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (argmax != nullptr) argmax[idx + i] = am[i];
            if (count_it[i]) {
                local += logf(sum[i]) + mx[i] - vl[i];             // -log softmax(label) from the logits (log-sum-exp form)
                counted += 1.f;
            }
        }
    }
    if (loss_out != nullptr) grid_sum2_store(local, counted, ws, loss_out, 1.f, true);
}

// dlogits = ce_scale * (p - onehot(label)) + p * (dp - sum_c p*dp)
//   dloss: device scalar (upstream gradient of the MEAN cross-entropy), count: device scalar (number of counted
//          pixels, loss_out[1] of the forward), or null;   dprobs: upstream gradient w.r.t. the probabilities, or null
template <int V>
__global__ void __launch_bounds__(256) seg_head_bwd_kernel(const float* __restrict__ probs,
                                                           const long long* __restrict__ labels,
                                                           const float* __restrict__ dloss,
                                                           const float* __restrict__ count,
                                                           const float* __restrict__ dprobs, int N, int C,
                                                           long long HW, float* __restrict__ dlogits) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)N * HW / V;
    const float ce = (dloss != nullptr && labels != nullptr) ? (*dloss) / (*count) : 0.f;
    for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < total;
         gidx += (long long)gridDim.x * blockDim.x) {
        const long long idx = gidx * V;
        const long long n = idx / HW, pix = idx - n * HW;
        const long long base = n * C * HW + pix;
        // two passes over the class planes (the second one hits the caches): holding p and dp of 4 pixels x 21 classes
        // in registers needs 255 of them plus spills
        float dot[V];
#pragma unroll
        for (int i = 0; i < V; ++i) dot[i] = 0.f;
        if (dprobs != nullptr) {
#pragma unroll 7
            for (int c = 0; c < C; ++c) {
                float p[V], dp[V];
                ldv<V>(probs + base + c * HW, p);
                ldv<V>(dprobs + base + c * HW, dp);
#pragma unroll
                for (int i = 0; i < V; ++i) dot[i] += p[i] * dp[i];
            }
        }
        long long lab[V];
        float cew[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            lab[i] = labels != nullptr ? labels[idx + i] : -1;
            cew[i] = (lab[i] >= 0 && lab[i] < C) ? ce : 0.f;        // ignored pixels carry no cross-entropy gradient
        }
#pragma unroll 7
        for (int c = 0; c < C; ++c) {
            {
                float p[V], dp[V], g[V];
                ldv<V>(probs + base + c * HW, p);
                if (dprobs != nullptr) {
                    ldv<V>(dprobs + base + c * HW, dp);
                } else {
#pragma unroll
                    for (int i = 0; i < V; ++i) dp[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    g[i] = p[i] * (dp[i] - dot[i]);
                    g[i] += cew[i] * (p[i] - (c == (int)lab[i] ? 1.f : 0.f));
                }
                stv<V>(dlogits + base + c * HW, g);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// LSGAN patch loss (nn.MSELoss against an all-ones / all-zeros target, model.py:270,445-446,452,
// 521-534) and L1 loss (nn.L1Loss, model.py:271,453,461), forward sum and gradient seed in one
// elementwise pass each.  The reference materialises the target tensor, the difference, its square
// (or abs) and a mean: 4-5 ATen kernels forward and 3 backward per loss; here the target is a scalar.
// ---------------------------------------------------------------------------------------------
// loss_out = scale * sum (x - target)^2
__global__ void __launch_bounds__(256) lsgan_fwd_kernel(const float* __restrict__ x, long long n, float target,
                                                        float scale, float* __restrict__ loss_sum, float* __restrict__ ws) {
    pdl_wait();
    pdl_launch();
    float local = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = x[i] - target;
        local += d * d;
    }
    grid_sum2_store(local, 0.f, ws, loss_sum, scale, false);
}
// dx = dloss * 2 (x - target) / n      (dloss: device scalar, gradient of the MEAN)
__global__ void __launch_bounds__(256) lsgan_bwd_kernel(const float* __restrict__ x, long long n, float target,
                                                        const float* __restrict__ dloss, float* __restrict__ dx) {
    pdl_wait();
    pdl_launch();
    const float sc = 2.f * (*dloss) / (float)n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = sc * (x[i] - target);
}
// loss_out = scale * sum |x - y|.  VEC: both pointers are 16-byte aligned (128-bit loads); otherwise scalar loads
// (nn.L1Loss takes any contiguous slice, e.g. l_img[rank * per : (rank + 1) * per]).
template <bool VEC>
__global__ void __launch_bounds__(256) l1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     long long n, float scale, float* __restrict__ loss_sum,
                                                     float* __restrict__ ws) {
    pdl_wait();
    pdl_launch();
    float local = 0.f;
    if (VEC) {
        const long long n4 = n >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const float4* y4 = reinterpret_cast<const float4*>(y);
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            const float4 a = x4[i], b = y4[i];
            local += fabsf(a.x - b.x) + fabsf(a.y - b.y) + fabsf(a.z - b.z) + fabsf(a.w - b.w);
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) local += fabsf(x[n4 * 4 + threadIdx.x] - y[n4 * 4 + threadIdx.x]);
    } else {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            local += fabsf(x[i] - y[i]);
    }
    grid_sum2_store(local, 0.f, ws, loss_sum, scale, false);
}
// dx = dloss * sign(x - y) / n   (sign(0) = 0, as torch)
__global__ void __launch_bounds__(256) l1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     long long n, const float* __restrict__ dloss,
                                                     float* __restrict__ dx) {
    pdl_wait();
    pdl_launch();
    const float sc = (*dloss) / (float)n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = x[i] - y[i];
        dx[i] = d > 0.f ? sc : (d < 0.f ? -sc : 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// Adam over one flat fp32 bucket (torch.optim.Adam semantics without weight decay / amsgrad,
// model.py:286-287,474,542): exp_avg, exp_avg_sq, bias corrections from a device-side step
// counter (already incremented for this step) and a device-side learning rate, so the update is
// CUDA-graph capturable and LambdaLR (utils.py:434-441, model.py:289-290,659-660) only rewrites
// one float.  One launch per optimizer instead of a multi-tensor-apply pass over ~100 tensors.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        const float* __restrict__ lr, float beta1, float beta2,
                                                        float eps, const float* __restrict__ step) {
    pdl_wait();
    pdl_launch();
    const float t = *step;
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2 = 1.f - powf(beta2, t);
    const float step_size = (*lr) / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        mm = mm + (gg - mm) * (1.f - beta1);
        vv = beta2 * vv + (1.f - beta2) * gg * gg;
        const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        pp = pp - step_size * (mm / denom);
    };
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = g4[i];
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = n4 * 4 + threadIdx.x;
        upd(p[i], g[i], m[i], v[i]);
    }
}


// ---------------------------------------------------------------------------------------------
// Confusion matrix of a label map against its prediction (utils.py:357-372, runningScore._fast_hist:
// np.bincount(n_class * true[mask] + pred[mask]) with mask = 0 <= true < n_class), accumulated on the device:
// per-block histogram in shared memory, one 64-bit atomic per non-empty bin and block.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) confusion_kernel(const long long* __restrict__ lt, const long long* __restrict__ lp,
                                                        long long n, int C, unsigned long long* __restrict__ hist) {
    pdl_wait();
    pdl_launch();
    __shared__ unsigned int s_hist[kMaxClasses * kMaxClasses];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long t = lt[i], p = lp[i];
        if (t >= 0 && t < C && p >= 0 && p < C) atomicAdd(&s_hist[(int)t * C + (int)p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}
}  // namespace sscg

using namespace sscg;

// measured at 16 / 32 x 21 x 256 x 256 (tools/bench_seg.py): forward 122 / 123 / 142 us with 1 / 2 / 4 pixels per thread,
// backward 204 / 199 / 137 us
constexpr int kSegVecFwd = 4, kSegVecBwd = 4;
static bool seg_vec_ok(int kSegVec, int64_t HW, const void* a, const void* b, const void* c) {
    if (kSegVec == 1 || HW % kSegVec) return false;
    const uintptr_t m = 4 * kSegVec - 1;
    return !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & m);
}

extern "C" int sscg_seg_head_fwd(const float* logits, const int64_t* labels, int32_t N, int32_t C, int64_t HW,
                                 int64_t ignore_index, float* probs, int64_t* argmax, float* loss_out, void* ws,
                                 void* stream) {
    if (C < 1 || C > kMaxClasses) return set_error("seg_head_fwd: C=%d must be in [1, %d]", C, kMaxClasses);
    if (loss_out != nullptr && ws == nullptr) return set_error("seg_head_fwd: the loss needs a workspace of SSCG_LOSS_WS_BYTES");
    const bool vec = seg_vec_ok(kSegVecFwd, HW, logits, probs, nullptr);
    const long long total = (long long)N * HW / (vec ? kSegVecFwd : 1);
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        if (vec)
            launch_k(seg_head_fwd_kernel<kSegVecFwd>, (int)g, 256, 0, static_cast<cudaStream_t>(stream),
                logits, reinterpret_cast<const long long*>(labels), N, C, HW, (long long)ignore_index, probs,
                reinterpret_cast<long long*>(argmax), loss_out, reinterpret_cast<float*>(ws));
        else
            launch_k(seg_head_fwd_kernel<1>, (int)g, 256, 0, static_cast<cudaStream_t>(stream),
                logits, reinterpret_cast<const long long*>(labels), N, C, HW, (long long)ignore_index, probs,
                reinterpret_cast<long long*>(argmax), loss_out, reinterpret_cast<float*>(ws));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("seg_head_fwd launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int sscg_seg_head_bwd(const float* probs, const int64_t* labels, const float* dloss, const float* count,
                                 const float* dprobs, int32_t N, int32_t C, int64_t HW, float* dlogits, void* stream) {
    if (C < 1 || C > kMaxClasses) return set_error("seg_head_bwd: C=%d must be in [1, %d]", C, kMaxClasses);
    if (dloss != nullptr && labels != nullptr && count == nullptr) return set_error("seg_head_bwd: dloss needs the pixel count");
    const bool vec = seg_vec_ok(kSegVecBwd, HW, probs, dprobs, dlogits);
    const long long total = (long long)N * HW / (vec ? kSegVecBwd : 1);
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        if (vec)
            launch_k(seg_head_bwd_kernel<kSegVecBwd>, (int)g, 256, 0, static_cast<cudaStream_t>(stream),
                probs, reinterpret_cast<const long long*>(labels), dloss, count, dprobs, N, C, HW, dlogits);
        else
            launch_k(seg_head_bwd_kernel<1>, (int)g, 256, 0, static_cast<cudaStream_t>(stream),
                probs, reinterpret_cast<const long long*>(labels), dloss, count, dprobs, N, C, HW, dlogits);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("seg_head_bwd launch: %s", cudaGetErrorString(e));
    return 0;
}

static inline int loss_grid(long long n) {
    long long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    return (int)g;
}
#define SSCG_LOSS_LAUNCH_CHECK(name)                                                         \
    do {                                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                 \
        if (e_ != cudaSuccess) return set_error(name " launch: %s", cudaGetErrorString(e_)); \
    } while (0)

extern "C" int sscg_lsgan_fwd(const float* x, int64_t n, float target, float scale, float* loss_sum, void* ws,
                              void* stream) {
    if (!x || !loss_sum || !ws || n < 1) return set_error("lsgan_fwd: bad arguments");
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        launch_k(lsgan_fwd_kernel, loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream), x, n, target, scale, loss_sum,
                                                                                     reinterpret_cast<float*>(ws));
    }
    SSCG_LOSS_LAUNCH_CHECK("lsgan_fwd");
    return 0;
}
extern "C" int sscg_lsgan_bwd(const float* x, int64_t n, float target, const float* dloss, float* dx, void* stream) {
    if (!x || !dloss || !dx || n < 1) return set_error("lsgan_bwd: bad arguments");
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        launch_k(lsgan_bwd_kernel, loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream), x, n, target, dloss, dx);
    }
    SSCG_LOSS_LAUNCH_CHECK("lsgan_bwd");
    return 0;
}
extern "C" int sscg_l1_fwd(const float* x, const float* y, int64_t n, float scale, float* loss_sum, void* ws,
                           void* stream) {
    if (!x || !y || !loss_sum || !ws || n < 1) return set_error("l1_fwd: bad arguments");
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        if (vec)
            launch_k(l1_fwd_kernel<true>, loss_grid(n >> 2), 256, 0, static_cast<cudaStream_t>(stream), 
                x, y, n, scale, loss_sum, reinterpret_cast<float*>(ws));
        else
            launch_k(l1_fwd_kernel<false>, loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream), 
                x, y, n, scale, loss_sum, reinterpret_cast<float*>(ws));
    }
    SSCG_LOSS_LAUNCH_CHECK("l1_fwd");
    return 0;
}
extern "C" int sscg_l1_bwd(const float* x, const float* y, int64_t n, const float* dloss, float* dx, void* stream) {
    if (!x || !y || !dloss || !dx || n < 1) return set_error("l1_bwd: bad arguments");
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        launch_k(l1_bwd_kernel, loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream), x, y, n, dloss, dx);
    }
    SSCG_LOSS_LAUNCH_CHECK("l1_bwd");
    return 0;
}
extern "C" int sscg_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float beta1,
                              float beta2, float eps, const float* step, void* stream) {
    if (!p || !g || !m || !v || !lr || !step || n < 1) return set_error("adam_flat: bad arguments");
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
         reinterpret_cast<uintptr_t>(v)) & 15)
        return set_error("adam_flat: buffers must be 16-byte aligned");
    {
        LaunchScope ls_(11, static_cast<cudaStream_t>(stream));
        launch_k(adam_flat_kernel, loss_grid(n >> 2), 256, 0, static_cast<cudaStream_t>(stream), p, g, m, v, n, lr, beta1, beta2,
                                                                                         eps, step);
    }
    SSCG_LOSS_LAUNCH_CHECK("adam_flat");
    return 0;
}

extern "C" int sscg_confusion(const int64_t* label_true, const int64_t* label_pred, int64_t n, int32_t n_class,
                              uint64_t* hist, void* stream) {
    if (!label_true || !label_pred || !hist || n < 1) return set_error("confusion: bad arguments");
    if (n_class < 1 || n_class > kMaxClasses) return set_error("confusion: n_class=%d must be in [1, %d]", n_class, kMaxClasses);
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        launch_k(confusion_kernel, loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream), 
            reinterpret_cast<const long long*>(label_true), reinterpret_cast<const long long*>(label_pred), n, n_class,
            reinterpret_cast<unsigned long long*>(hist));
    }
    SSCG_LOSS_LAUNCH_CHECK("confusion");
    return 0;
}
