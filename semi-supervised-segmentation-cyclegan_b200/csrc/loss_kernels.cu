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
