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

// logits [N][C][H][W] fp32; labels [N][H][W] int64 (or null); probs [N][C][H][W] (or null);
// argmax [N][H][W] int64 (or null); loss_sum: fp32 accumulator of -log p[label] (or null)
__global__ void __launch_bounds__(256) seg_head_fwd_kernel(const float* __restrict__ logits,
                                                           const long long* __restrict__ labels, int N, int C,
                                                           long long HW, float* __restrict__ probs,
                                                           long long* __restrict__ argmax,
                                                           float* __restrict__ loss_sum) {
    const long long total = (long long)N * HW;
    float local = 0.f;
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
        if (labels != nullptr && loss_sum != nullptr) {
            const int lab = (int)labels[idx];
            float pl = 0.f;
#pragma unroll
            for (int c = 0; c < kMaxClasses; ++c)
                if (c == lab) pl = v[c];
            local += -__logf(fmaxf(pl * inv, 1e-38f));
        }
    }
    if (loss_sum != nullptr) {
        for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
        __shared__ float s[8];
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < (blockDim.x >> 5); ++i) t += s[i];
            atomicAdd(loss_sum, t);
        }
    }
}

// dlogits = ce_scale * (p - onehot(label)) + p * (dp - sum_c p*dp)
//   ce_scale: device scalar pointer (upstream gradient of the MEAN cross-entropy) times 1/(N*HW), or null
//   dprobs:   upstream gradient w.r.t. the probabilities, or null
__global__ void __launch_bounds__(256) seg_head_bwd_kernel(const float* __restrict__ probs,
                                                           const long long* __restrict__ labels,
                                                           const float* __restrict__ dloss,
                                                           const float* __restrict__ dprobs, int N, int C,
                                                           long long HW, float* __restrict__ dlogits) {
    const long long total = (long long)N * HW;
    const float ce = (dloss != nullptr && labels != nullptr) ? (*dloss) / (float)total : 0.f;
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
        const int lab = labels != nullptr ? (int)labels[idx] : -1;
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c) {
            if (c < C) {
                float g = p[c] * (dp[c] - dot);
                g += ce * (p[c] - (c == lab ? 1.f : 0.f));
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
__device__ __forceinline__ void block_sum_atomic(float local, float* dst, float scale) {
    for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    __shared__ float s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) t += s[i];
        atomicAdd(dst, t * scale);
    }
}

// loss_sum += scale * sum (x - target)^2
__global__ void __launch_bounds__(256) lsgan_fwd_kernel(const float* __restrict__ x, long long n, float target,
                                                        float scale, float* __restrict__ loss_sum) {
    float local = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = x[i] - target;
        local += d * d;
    }
    block_sum_atomic(local, loss_sum, scale);
}
// dx = dloss * 2 (x - target) / n      (dloss: device scalar, gradient of the MEAN)
__global__ void __launch_bounds__(256) lsgan_bwd_kernel(const float* __restrict__ x, long long n, float target,
                                                        const float* __restrict__ dloss, float* __restrict__ dx) {
    const float sc = 2.f * (*dloss) / (float)n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = sc * (x[i] - target);
}
// loss_sum += scale * sum |x - y|
__global__ void __launch_bounds__(256) l1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     long long n, float scale, float* __restrict__ loss_sum) {
    float local = 0.f;
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = x4[i], b = y4[i];
        local += fabsf(a.x - b.x) + fabsf(a.y - b.y) + fabsf(a.z - b.z) + fabsf(a.w - b.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) local += fabsf(x[n4 * 4 + threadIdx.x] - y[n4 * 4 + threadIdx.x]);
    block_sum_atomic(local, loss_sum, scale);
}
// dx = dloss * sign(x - y) / n   (sign(0) = 0, as torch)
__global__ void __launch_bounds__(256) l1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     long long n, const float* __restrict__ dloss,
                                                     float* __restrict__ dx) {
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
                                 float* probs, int64_t* argmax, float* loss_sum, void* stream) {
    if (C < 1 || C > kMaxClasses) return set_error("seg_head_fwd: C=%d must be in [1, %d]", C, kMaxClasses);
    const long long total = (long long)N * HW;
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        seg_head_fwd_kernel<<<(int)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            logits, reinterpret_cast<const long long*>(labels), N, C, HW, probs, reinterpret_cast<long long*>(argmax),
            loss_sum);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("seg_head_fwd launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int sscg_seg_head_bwd(const float* probs, const int64_t* labels, const float* dloss, const float* dprobs,
                                 int32_t N, int32_t C, int64_t HW, float* dlogits, void* stream) {
    if (C < 1 || C > kMaxClasses) return set_error("seg_head_bwd: C=%d must be in [1, %d]", C, kMaxClasses);
    const long long total = (long long)N * HW;
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        seg_head_bwd_kernel<<<(int)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            probs, reinterpret_cast<const long long*>(labels), dloss, dprobs, N, C, HW, dlogits);
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

extern "C" int sscg_lsgan_fwd(const float* x, int64_t n, float target, float scale, float* loss_sum, void* stream) {
    if (!x || !loss_sum || n < 1) return set_error("lsgan_fwd: bad arguments");
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        lsgan_fwd_kernel<<<loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, target, scale, loss_sum);
    }
    SSCG_LOSS_LAUNCH_CHECK("lsgan_fwd");
    return 0;
}
extern "C" int sscg_lsgan_bwd(const float* x, int64_t n, float target, const float* dloss, float* dx, void* stream) {
    if (!x || !dloss || !dx || n < 1) return set_error("lsgan_bwd: bad arguments");
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        lsgan_bwd_kernel<<<loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, target, dloss, dx);
    }
    SSCG_LOSS_LAUNCH_CHECK("lsgan_bwd");
    return 0;
}
extern "C" int sscg_l1_fwd(const float* x, const float* y, int64_t n, float scale, float* loss_sum, void* stream) {
    if (!x || !y || !loss_sum || n < 1) return set_error("l1_fwd: bad arguments");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return set_error("l1_fwd: pointers must be 16-byte aligned");
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        l1_fwd_kernel<<<loss_grid(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, scale, loss_sum);
    }
    SSCG_LOSS_LAUNCH_CHECK("l1_fwd");
    return 0;
}
extern "C" int sscg_l1_bwd(const float* x, const float* y, int64_t n, const float* dloss, float* dx, void* stream) {
    if (!x || !y || !dloss || !dx || n < 1) return set_error("l1_bwd: bad arguments");
    {
        LaunchScope ls_(10, static_cast<cudaStream_t>(stream));
        l1_bwd_kernel<<<loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, dloss, dx);
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
        adam_flat_kernel<<<loss_grid(n >> 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2,
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
        confusion_kernel<<<loss_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const long long*>(label_true), reinterpret_cast<const long long*>(label_pred), n, n_class,
            reinterpret_cast<unsigned long long*>(hist));
    }
    SSCG_LOSS_LAUNCH_CHECK("confusion");
    return 0;
}
