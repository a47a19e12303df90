// sscg_common.cuh — host-side helpers shared by the C-ABI translation units: error string,
// tensor-map (TMA descriptor) encoding through the lazily resolved driver entry point.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/sscg_b200.h"
#include "sscg_ptx.cuh"

namespace sscg {

int set_error(const char* fmt, ...);   // formats into the last-error buffer, returns 1

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// Launch accounting + optional CUDA-event bracketing per profiling tag (see sscg_prof_begin).
struct LaunchScope {
    int slot;
    cudaStream_t stream;
    LaunchScope(int tag, cudaStream_t s);
    ~LaunchScope();
};

// Every kernel of the library is launched through launch_k: with programmatic dependent launch (PDL) enabled the launch
// may be SCHEDULED while its predecessor in the stream is still draining — CTAs become resident as SMs free up, run
// their prologue (barrier init, TMEM allocation, descriptor prefetch) and block in pdl_wait() until the predecessor
// has completed and its writes are visible.  Every kernel therefore calls pdl_wait() before its first access to
// global memory (reads AND writes: the predecessor may still be reading what this kernel overwrites), followed by
// pdl_launch() so that its own successor can be scheduled.  Inside CUDA graphs the attribute becomes a programmatic
// dependency edge.  OFF by default (SSCG_PDL=1 / sscg_set_pdl(1) turn it on): on the power-capped B200s of this pool
// the step did not get faster (47.99 / 48.21 ms without, 48.44 / 48.38 ms with: the SM clock drops with the idle gaps).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// 4-D map over an NHWC bf16 view: dims (C, W, H, N), 128B swizzle, zero fill outside the view.
// swizzle_bytes: 128 (default), 64, 32 or 0 (dense rows)
int encode_view_4d(CUtensorMap* tm, const SscgView& v, const void* ptr, const uint32_t box[4], const uint32_t es[4],
                   int swizzle_bytes = 128);
// 2-D map over a row-major bf16 matrix [rows][cols]
int encode_2d(CUtensorMap* tm, const void* ptr, int cols, int rows, int box_cols, int box_rows);

}  // namespace sscg
