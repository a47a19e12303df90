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

// 4-D map over an NHWC bf16 view: dims (C, W, H, N), 128B swizzle, zero fill outside the view.
// swizzle_bytes: 128 (default), 64, 32 or 0 (dense rows)
int encode_view_4d(CUtensorMap* tm, const SscgView& v, const void* ptr, const uint32_t box[4], const uint32_t es[4],
                   int swizzle_bytes = 128);
// 2-D map over a row-major bf16 matrix [rows][cols]
int encode_2d(CUtensorMap* tm, const void* ptr, int cols, int rows, int box_cols, int box_rows);

}  // namespace sscg
