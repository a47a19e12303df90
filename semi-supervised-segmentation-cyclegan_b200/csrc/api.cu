// api.cu — unity translation unit of libsscg_b200.so: shared host helpers + all kernels.
// Built by csrc/build.sh (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo).
#include "sscg_common.cuh"

namespace sscg {

static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

int encode_view_4d(CUtensorMap* tm, const SscgView& v, const void* ptr, const uint32_t box[4], const uint32_t es[4],
                   int swizzle_bytes) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sN * 2};
    cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t s[4] = {es[0], es[1], es[2], es[3]};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
        return set_error("tensor map: view pointer/strides must be 16-byte aligned (ptr=%p sW=%lld sH=%lld sN=%lld)", ptr,
                         (long long)v.sW, (long long)v.sH, (long long)v.sN);
    const CUtensorMapSwizzle sw = swizzle_bytes == 0    ? CU_TENSOR_MAP_SWIZZLE_NONE
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                        : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, b, s,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error("cuTensorMapEncodeTiled(4d) failed: %d dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) "
                         "box=(%u,%u,%u,%u) es=(%u,%u,%u,%u)",
                         (int)r, dims[0], dims[1], dims[2], dims[3], strides[0], strides[1], strides[2], b[0], b[1], b[2],
                         b[3], s[0], s[1], s[2], s[3]);
    return 0;
}

int encode_2d(CUtensorMap* tm, const void* ptr, int cols, int rows, int box_cols, int box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t b[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t s[2] = {1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15))
        return set_error("tensor map: matrix pointer/pitch must be 16-byte aligned");
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, b, s,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error("cuTensorMapEncodeTiled(2d) failed: %d dims=(%llu,%llu) box=(%u,%u)", (int)r, dims[0], dims[1],
                         b[0], b[1]);
    return 0;
}

// ---- launch accounting / per-tag event timing ---------------------------------------------------
static unsigned long long g_launches = 0;
constexpr int kProfMax = 8192;
static struct {
    bool on = false;
    bool created = false;
    unsigned mask = 0xffffu;
    int n = 0;
    cudaEvent_t ev[2 * kProfMax];
    int tag[kProfMax];
} g_prof;

static int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("SSCG_PDL");
        g_pdl = (e && e[0] == '1') ? 1 : 0;     // measured: no gain under the power cap -> opt-in
    }
    return g_pdl != 0;
}
void pdl_set(int on) { g_pdl = on ? 1 : 0; }

LaunchScope::LaunchScope(int tag, cudaStream_t s) : slot(-1), stream(s) {
    ++g_launches;
    if (g_prof.on && ((g_prof.mask >> (tag & 15)) & 1u) && g_prof.n < kProfMax) {
        slot = g_prof.n++;
        g_prof.tag[slot] = tag & 15;
        cudaEventRecord(g_prof.ev[2 * slot], stream);
    }
}
LaunchScope::~LaunchScope() {
    if (slot >= 0) cudaEventRecord(g_prof.ev[2 * slot + 1], stream);
}

}  // namespace sscg

extern "C" uint64_t sscg_launch_count(void) { return sscg::g_launches; }
extern "C" int sscg_set_pdl(int32_t on) { sscg::pdl_set(on); return 0; }

extern "C" int sscg_prof_begin(uint32_t tag_mask) {
    using namespace sscg;
    g_prof.mask = tag_mask ? tag_mask : 0xffffu;
    if (!g_prof.created) {
        for (int i = 0; i < 2 * kProfMax; ++i)
            if (cudaEventCreate(&g_prof.ev[i]) != cudaSuccess) return set_error("prof_begin: cudaEventCreate failed");
        g_prof.created = true;
    }
    g_prof.n = 0;
    g_prof.on = true;
    return 0;
}

extern "C" int sscg_prof_end(float* sum_ms, int32_t* count) {
    using namespace sscg;
    g_prof.on = false;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return set_error("prof_end: %s", cudaGetErrorString(e));
    for (int t = 0; t < 16; ++t) { sum_ms[t] = 0.f; count[t] = 0; }
    for (int i = 0; i < g_prof.n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]) == cudaSuccess) {
            sum_ms[g_prof.tag[i]] += ms;
            count[g_prof.tag[i]] += 1;
        }
    }
    return g_prof.n >= kProfMax ? 2 : 0;   // 2: event pool exhausted (partial coverage)
}

extern "C" const char* sscg_last_error(void) { return sscg::g_err; }
extern "C" int sscg_version(void) { return 100; }

extern "C" int sscg_device_error(void) {
    unsigned int v = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, sscg::g_sscg_dev_error, sizeof(v));
    if (e != cudaSuccess) return -1;
    if (v) {
        unsigned int z = 0;
        cudaMemcpyToSymbol(sscg::g_sscg_dev_error, &z, sizeof(z));
    }
    return (int)v;
}

extern "C" int sscg_fill_zero(void* ptr, int64_t bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return sscg::set_error("fill_zero: %s", cudaGetErrorString(e));
    return 0;
}

#include "conv_igemm.cu"
#include "conv_wgrad.cu"
#include "conv_wgrad7.cu"
#include "conv_nexp.cu"
#include "elementwise.cu"
#include "loss_kernels.cu"
