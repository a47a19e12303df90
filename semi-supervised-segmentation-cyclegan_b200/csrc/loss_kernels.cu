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
__global__ void __launch_bounds__(256) seg_head_fwd_kernel(const float* __restrict__ logits,
                                                           const long long* __restrict__ labels, int N, int C,
                                                           long long HW, long long ignore_index,
                                                           float* __restrict__ probs, long long* __restrict__ argmax,
                                                           float* __restrict__ loss_out, float* __restrict__ ws) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)N * HW;
    float local = 0.f, counted = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / HW, pix = idx - n * HW;
        const float* src = logits + n * C * HW + pix;
        float v[kMaxClasses];
        float mx = -INFINITY;
        int am = 0;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
            if (c < C) {
                v[c] = src[c * HW];
                if (v[c] > mx) { mx = v[c]; am = c; }      // strict '>' keeps the FIRST maximum (torch.max rule)
            }
        }
        float vl = 0.f;
        bool count_it = false;
        if (labels != nullptr && loss_out != nullptr) {
            const long long lab = labels[idx];
            if (lab >= 0 && lab < C) {
                count_it = true;
#pragma unroll
                for (int c = 0; c < kMaxClasses; ++c)
                    if (c == (int)lab) vl = v[c];
            } else if (lab != ignore_index) {
                atomicCAS(&g_sscg_dev_error, 0u, (31u << 16) | 0x80000000u);     // label outside [0, C)
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
            if (c < C) {
                v[c] = __expf(v[c] - mx);
                sum += v[c];
            }
        }
        const float inv = 1.f / sum;
        if (probs != nullptr) {
            float* dst = probs + n * C * HW + pix;
#pragma unroll
            for (int c = 0; c < kMaxClasses; ++c)
                if (c < C) dst[c * HW] = v[c] * inv;
        }
        if (argmax != nullptr) argmax[idx] = am;
        if (count_it) {
            local += logf(sum) + mx - vl;                  // -log softmax(label) from the logits (log-sum-exp form)
            counted += 1.f;
        }
    }
    if (loss_out != nullptr) grid_sum2_store(local, counted, ws, loss_out, 1.f, true);
}

// dlogits = ce_scale * (p - onehot(label)) + p * (dp - sum_c p*dp)
//   dloss: device scalar (upstream gradient of the MEAN cross-entropy), count: device scalar (number of counted
//          pixels, loss_out[1] of the forward), or null;   dprobs: upstream gradient w.r.t. the probabilities, or null
__global__ void __launch_bounds__(256) seg_head_bwd_kernel(const float* __restrict__ probs,
                                                           const long long* __restrict__ labels,
                                                           const float* __restrict__ dloss,
                                                           const float* __restrict__ count,
                                                           const float* __restrict__ dprobs, int N, int C,
                                                           long long HW, float* __restrict__ dlogits) {
    pdl_wait();
    pdl_launch();
    const long long total = (long long)N * HW;
    const float ce = (dloss != nullptr && labels != nullptr) ? (*dloss) / (*count) : 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long n = idx / HW, pix = idx - n * HW;
        const long long base = n * C * HW + pix;
        float p[kMaxClasses], dp[kMaxClasses];
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
            if (c < C) {
                p[c] = probs[base + c * HW];
                dp[c] = dprobs != nullptr ? dprobs[base + c * HW] : 0.f;
                dot += p[c] * dp[c];
            }
        }
        const long long lab = labels != nullptr ? labels[idx] : -1;
        const float cew = (lab >= 0 && lab < C) ? ce : 0.f;        // ignored pixels carry no cross-entropy gradient
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
            if (c < C) {
                float g = p[c] * (dp[c] - dot);
                g += cew * (p[c] - (c == (int)lab ? 1.f : 0.f));
                dlogits[base + c * HW] = g;
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

extern "C" int sscg_seg_head_fwd(const float* logits, const int64_t* labels, int32_t N, int32_t C, int64_t HW,
                                 int64_t ignore_index, float* probs, int64_t* argmax, float* loss_out, void* ws,
                                 void* stream) {
    if (C < 1 || C > kMaxClasses) return set_error("seg_head_fwd: C=%d must be in [1, %d]", C, kMaxClasses);
    if (loss_out != nullptr && ws == nullptr) return set_error("seg_head_fwd: the loss needs a workspace of SSCG_LOSS_WS_BYTES");
    const long long total = (long long)N * HW;
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        launch_k(seg_head_fwd_kernel, (int)g, 256, 0, static_cast<cudaStream_t>(stream), 
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
    const long long total = (long long)N * HW;
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        launch_k(seg_head_bwd_kernel, (int)g, 256, 0, static_cast<cudaStream_t>(stream), 
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
