// conv_wgrad.cu — weight gradient of the convolutions as a pixel-contraction GEMM on tcgen05.
//
//   dWt[tap][co][k] += sum over pixels  dY[pixel][co] * X[pixel*stride + tap][k]
//
// GEMM view: M = output channels (128 per CTA), N = BN columns of the K axis of the weight slab,
// contraction over pixels in blocks of 64.  Both operands are "MN-major" for the tensor core: a TMA
// box of TH x TW pixels x 64 channels lands as 64 rows (pixels = GEMM-K) of 128 B (64 channels =
// GEMM-M or -N), 128B-swizzled, which is exactly the canonical MN-major SWIZZLE_128B atom.
// Split-K over (sample, pixel tile) ranges.  The partial tiles of an output tile are combined in split order — every
// split CTA reduces its own slice of the tile once all partials are published (fixed-order reduction, sscg_ptx.cuh) —
// so the result does not depend on CTA scheduling (the earlier red.global.add.f32 version did).
//
// Replaces (reference): the cuDNN wgrad that autograd runs for every nn.Conv2d / ConvTranspose2d of
// arch/ops.py:40-57,63,68, arch/generators.py:74-90, arch/discriminators.py:45-58 during
// gen_loss.backward() / discriminator_loss.backward() (model.py:472,539).
#include "sscg_common.cuh"

namespace sscg {

struct WgradDev {
    int N;
    int tiles_h, tiles_w;   // pixel blocks per sample
    int TH, TW;
    int stride, org_h, org_w;
    SscgTap taps[SSCG_MAX_TAPS];
    int Co_pad, Kc, n_ktiles;
    int ksplit;
    float* dw;
    float* part;            // [tile][ksplit][128][BN] fp32 partial tiles (ksplit > 1)
    unsigned int* ctr;      // [tile][2] arrival / departure counters (zero between launches)
};

constexpr int kWgAtomBytes = 64 * 128;   // 64 pixels x 64 channels bf16

template <int BN, int SPLIT>
struct WgradCfg {
    static constexpr int kPlanes = (SPLIT == 3) ? 2 : 1;
    static constexpr int kAAtoms = 2;           // M = 128
    static constexpr int kBAtoms = BN / 64;
    static constexpr int kStageBytes = kPlanes * (kAAtoms + kBAtoms) * kWgAtomBytes;
    static constexpr int kMaxStages = (200 * 1024) / kStageBytes;
    // Two CTAs per SM wherever each still gets a ring of >= 2 stages: a CTA owns ONE output tile, so its epilogue
    // (TMEM -> partial tile, arrival wait, slice reduction: ~7 us of an 80 us launch) cannot overlap its own MMAs —
    // the co-resident CTAs' MMAs fill that time (80.5 -> 75.4 us on the residual-block shape, 300 -> 162 us on the
    // 64-channel stem, whose CTAs are short).
    // measured on the step's shapes: three CTAs pay off for the 64-wide tile only (138 vs 162 us on the stem); with
    // BN = 128 the third CTA leaves 2-stage rings and a 49-way split (84 vs 50 us), four are worse everywhere
    static constexpr int stages_for(int n) { return ((228 * 1024) / n - 1024 - 1280) / kStageBytes; }
    static constexpr int kCtasPerSm = BN > 256 ? 1
                                      : (BN == 64 && stages_for(3) >= 2) ? 3
                                      : (stages_for(2) >= 2) ? 2 : 1;
    static constexpr int kStages = kCtasPerSm == 1 ? (kMaxStages > 6 ? 6 : kMaxStages)
                                                   : (stages_for(kCtasPerSm) > 6 ? 6 : stages_for(kCtasPerSm));
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
    static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
    static constexpr int kN0 = BN > 256 ? 256 : BN;     // one tcgen05.mma covers at most N = 256 columns
    static constexpr int kN1 = BN - kN0;
    static_assert(kStages >= 2, "wgrad tile does not leave room for a 2-stage pipeline");
    static_assert(SPLIT == 1 || kN1 == 0, "wide wgrad tiles are bf16-mode only");
};

constexpr int kWgCtrs = 1024;           // arrival counters at the head of the workspace

// RW: pixel-row mode for stride-1 stems (bf16 mode, 1 x 64 pixel blocks, BN = 64 * G columns for 8 * G input channels):
// the X operand of a filter row is the image row itself.  Per channel group one dense box of 64 + 8 pixels x 8 channels
// (16 bytes per pixel) is read as an MN-major operand with OVERLAPPING rows — K row = pixel (16-byte pitch), the next
// 16-byte chunk along N = the next pixel (SBO 16), 8 K rows = 128 bytes (LBO; the no-swizzle MN-major roles of the two
// offsets are the reverse of the K-major ones) — instead of a TMA-expanded 8 KB window
// atom per group (see conv_igemm.cu, RW).  One N = 64 MMA per group and 16-pixel step.
template <int BN, int SPLIT, bool RW>
__global__ void __launch_bounds__(192, (WgradCfg<BN, SPLIT>::kCtasPerSm))
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmDyLo,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXLo,
                  const __grid_constant__ WgradDev p) {
    using Cfg = WgradCfg<BN, SPLIT>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kPlanes = Cfg::kPlanes;
    constexpr int kBAtoms = Cfg::kBAtoms;

    const int split = blockIdx.x;
    const int mt = blockIdx.y / p.n_ktiles;        // 128-wide output-channel tile
    const int kt = blockIdx.y - mt * p.n_ktiles;   // BN-wide K tile
    const SscgTap tap = p.taps[blockIdx.z];
    const int m0 = mt * 128;
    const int a_atoms = (p.Co_pad - m0) >= 128 ? 2 : 1;   // Co_pad is a multiple of 64

    const int blocks_per_sample = p.tiles_h * p.tiles_w;
    const int total_blocks = p.N * blocks_per_sample;
    const int per = (total_blocks + p.ksplit - 1) / p.ksplit;
    const int pb0 = split * per;
    const int pb1 = min(total_blocks, pb0 + per);
    if (pb0 >= pb1) return;
    const int nkb = pb1 - pb0;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full_bar = bars + 2 * kStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDy);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(tmem_full_bar), 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_ptr_smem), Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();          // the prologue above overlaps the predecessor's tail (launch_k, sscg_common.cuh)
    pdl_launch();

    // stage layout: [A hi atoms (2)] [B hi atoms] [A lo atoms (2)] [B lo atoms]
    constexpr int kAOff = 0;
    constexpr int kBOff = 2 * kWgAtomBytes;
    constexpr int kLoOff = (2 + kBAtoms) * kWgAtomBytes;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t par = 0;
            constexpr uint32_t kRwBox = (64 + 8) * 16;       // bytes of one pixel-row box
            const uint32_t tx = RW ? a_atoms * kWgAtomBytes + kBAtoms * kRwBox : kPlanes * (a_atoms + kBAtoms) * kWgAtomBytes;
            for (int kb = 0; kb < nkb; ++kb) {
                int b = pb0 + kb;
                const int n = b / blocks_per_sample; b -= n * blocks_per_sample;
                const int ti = b / p.tiles_w, tj = b - ti * p.tiles_w;
                const int i0 = ti * p.TH, j0 = tj * p.TW;
                mbar_wait(smem_u32(&empty_bar[stage]), par ^ 1, 4);
                const uint32_t fb = smem_u32(&full_bar[stage]);
                mbar_arrive_expect_tx(fb, tx);
                uint8_t* st = smem + stage * Cfg::kStageBytes;
                const int cw = j0 * p.stride + tap.dw + p.org_w;
                const int ch = i0 * p.stride + tap.dh + p.org_h;
                for (int a = 0; a < a_atoms; ++a)
                    tma_load_4d(smem_u32(st + kAOff + a * kWgAtomBytes), &tmDy, fb, m0 + a * 64, j0, i0, n);
#pragma unroll
                for (int q = 0; q < kBAtoms; ++q)
                    tma_load_4d(smem_u32(st + kBOff + q * kWgAtomBytes), &tmX, fb, RW ? q * 8 : kt * BN + q * 64, cw, ch, n);
                if (SPLIT == 3) {
                    for (int a = 0; a < a_atoms; ++a)
                        tma_load_4d(smem_u32(st + kLoOff + kAOff + a * kWgAtomBytes), &tmDyLo, fb, m0 + a * 64, j0, i0, n);
#pragma unroll
                    for (int q = 0; q < kBAtoms; ++q)
                        tma_load_4d(smem_u32(st + kLoOff + kBOff + q * kWgAtomBytes), &tmXLo, fb, kt * BN + q * 64, cw, ch, n);
                }
                if (++stage == kStages) { stage = 0; par ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, Cfg::kN0, 1, 1);
            constexpr uint32_t idesc1 = make_idesc_bf16(128, Cfg::kN1 > 0 ? Cfg::kN1 : 16, 1, 1);
            int stage = 0; uint32_t par = 0;
            uint32_t acc = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(smem_u32(&full_bar[stage]), par, 5);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + stage * Cfg::kStageBytes);
                // MN-major: LBO = stride between 64-wide channel atoms, SBO = stride between 8-pixel groups
                const uint64_t da = make_smem_desc_sw128(st + kAOff, kWgAtomBytes, 1024);
                const uint64_t db = make_smem_desc_sw128(st + kBOff, kWgAtomBytes, 1024);
                if (RW) {
                    constexpr uint32_t idesc64 = make_idesc_bf16(128, 64, 1, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int q = 0; q < kBAtoms; ++q) {      // group q: columns 64 q .., rows of 16 bytes, +16 pixels = +256 B
                            const uint64_t dq = make_smem_desc(st + kBOff + q * kWgAtomBytes, 128, 16, 0);   // MN-major: LBO = 8 K rows, SBO = next 16-byte chunk along N
                            umma_bf16(tmem_base + 64 * q, da + 128 * k, dq + 16 * k, idesc64, acc);
                        }
                        acc = 1;
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));
                    if (++stage == kStages) { stage = 0; par ^= 1; }
                    continue;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // 16 pixels per MMA: +2048 B
                    umma_bf16(tmem_base, da + 128 * k, db + 128 * k, idesc, acc);
                    if (Cfg::kN1 > 0) {   // columns 256.. of a wide tile (second group of 64-channel atoms)
                        const uint64_t db1 = make_smem_desc_sw128(st + kBOff + 4 * kWgAtomBytes, kWgAtomBytes, 1024);
                        umma_bf16(tmem_base + 256, da + 128 * k, db1 + 128 * k, idesc1, acc);
                    }
                    acc = 1;
                    if (SPLIT == 3) {
                        const uint64_t dalo = make_smem_desc_sw128(st + kLoOff + kAOff, kWgAtomBytes, 1024);
                        const uint64_t dblo = make_smem_desc_sw128(st + kLoOff + kBOff, kWgAtomBytes, 1024);
                        umma_bf16(tmem_base, dalo + 128 * k, db + 128 * k, idesc, 1);
                        umma_bf16(tmem_base, da + 128 * k, dblo + 128 * k, idesc, 1);
                    }
                }
                umma_commit(smem_u32(&empty_bar[stage]));
                if (++stage == kStages) { stage = 0; par ^= 1; }
            }
            umma_commit(smem_u32(tmem_full_bar));
        }
    } else {
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        const int co = m0 + m;
        const bool valid = (m < a_atoms * 64);
        mbar_wait(smem_u32(tmem_full_bar), 0, 6);
        tc_fence_after();
        float* drow = p.dw + ((long long)tap.brow * p.Co_pad + co) * p.Kc + kt * BN;
        const int nsplit = (total_blocks + per - 1) / per;          // splits that own at least one pixel block
        const int tile_id = blockIdx.z * gridDim.y + blockIdx.y;
        float* prow = p.part + (((long long)tile_id * p.ksplit + split) * 128 + m) * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, r);
            tmem_ld_wait();
            if (valid) {
                float4* dst = reinterpret_cast<float4*>((nsplit == 1 ? drow : prow) + c * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                           __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                    if (nsplit == 1) {          // sole owner of the tile: plain read-modify-write
                        const float4 o = dst[q];
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    dst[q] = v;
                }
            }
        }
        tc_fence_before();
        if (nsplit > 1) {
            // All splits of a tile are consecutive blocks (blockIdx.x fastest), hence co-resident or next in line for
            // an SM: each one publishes its partial tile, waits until all nsplit are there, and then reduces ITS slice
            // of the tile over the splits in split order — a fixed-order sum whose cost is spread over the nsplit SMs
            // (one CTA reducing the whole tile reads nsplit x 128 KB through a single SM: +60 us on the 3x3 layers).
            const int e = threadIdx.x - 64;
            unsigned int* ctr_a = p.ctr + 2 * tile_id;       // arrivals
            unsigned int* ctr_d = ctr_a + 1;                 // departures (the last one re-arms both counters)
            named_bar_sync(1, 128);                          // the CTA's partial tile is stored
            if (e == 0) {
                atom_add_acq_rel_gpu(ctr_a, 1u);
                spin_until_ge(ctr_a, (unsigned int)nsplit, 8);
            }
            named_bar_sync(2, 128);
            const float* pt = p.part + (long long)tile_id * p.ksplit * 128 * BN;
            const int rows = a_atoms * 64;
            float* dbase = p.dw + ((long long)tap.brow * p.Co_pad + m0) * p.Kc + kt * BN;
            const int total4 = rows * BN / 4;
            const int per4 = (total4 + nsplit - 1) / nsplit;
            const int j0 = split * per4, j1 = min(total4, j0 + per4);
            // two positions per thread and pass, all their partials (up to 8 splits each) in flight together: the loads
            // come from L2 (~1 us round trip), so memory-level parallelism is what bounds this loop
            for (int j = j0 + e; j < j1; j += 256) {
                int idx[2];
                bool on[2];
                float4 acc[2], old[2];
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int jw = j + w * 128;
                    on[w] = jw < j1;
                    idx[w] = (on[w] ? jw : j) * 4;
                    acc[w] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float4* d4[2];
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int row = idx[w] / BN, col = idx[w] - row * BN;
                    d4[w] = reinterpret_cast<float4*>(dbase + (long long)row * p.Kc + col);
                    old[w] = *d4[w];
                }
                for (int s0 = 0; s0 < nsplit; s0 += 8) {
                    float4 v[2][8];
#pragma unroll
                    for (int w = 0; w < 2; ++w)
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            v[w][u] = (s0 + u < nsplit) ? ld_cg_f4(pt + (long long)(s0 + u) * 128 * BN + idx[w])
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w = 0; w < 2; ++w)
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            acc[w].x += v[w][u].x; acc[w].y += v[w][u].y; acc[w].z += v[w][u].z; acc[w].w += v[w][u].w;
                        }
                }
#pragma unroll
                for (int w = 0; w < 2; ++w)
                    if (on[w])
                        *d4[w] = make_float4(old[w].x + acc[w].x, old[w].y + acc[w].y, old[w].z + acc[w].z, old[w].w + acc[w].w);
            }
            named_bar_sync(1, 128);                          // this CTA no longer reads the partial tiles
            if (e == 0) {
                const unsigned int old = atom_add_acq_rel_gpu(ctr_d, 1u);
                if (old + 1u == (unsigned int)nsplit) {      // everyone is past the wait: counters back to zero
                    *ctr_a = 0u;
                    *ctr_d = 0u;
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

template <int BN, int SPLIT, bool RW = false>
static int launch_wgrad(const CUtensorMap& tmDy, const CUtensorMap& tmDyLo, const CUtensorMap& tmX,
                        const CUtensorMap& tmXLo, const WgradDev& d, dim3 grid, cudaStream_t stream, int tag) {
    using Cfg = WgradCfg<BN, SPLIT>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<BN, SPLIT, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmemBytes);
        if (e != cudaSuccess) return set_error("conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    {
        LaunchScope ls(tag, stream);
        launch_k(conv_wgrad_kernel<BN, SPLIT, RW>, grid, 192, Cfg::kSmemBytes, stream, tmDy, tmDyLo, tmX, tmXLo, d);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("conv_wgrad<%d,%d> launch: %s", BN, SPLIT, cudaGetErrorString(e));
    return 0;
}

}  // namespace sscg

using namespace sscg;

extern "C" int32_t sscg_conv_wgrad_ctas_per_sm(int32_t BN, int32_t split) {
#define SSCG_WG(BN_) \
    case BN_: return split == 3 ? WgradCfg<BN_, 3>::kCtasPerSm : WgradCfg<BN_, 1>::kCtasPerSm;
    switch (BN) {
        SSCG_WG(64)
        SSCG_WG(128)
        SSCG_WG(256)
        case 192: return WgradCfg<192, 1>::kCtasPerSm;
        case 448: return WgradCfg<448, 1>::kCtasPerSm;
        default: return 1;
    }
#undef SSCG_WG
}

extern "C" int64_t sscg_conv_wgrad_ws_bytes(const SscgWgradArgs* a) {
    if (a->BN <= 0 || a->Kc % a->BN) return -1;
    if (a->ksplit <= 1) return 0;
    const long long tiles = (long long)((a->Co_pad + 127) / 128) * (a->Kc / a->BN) * a->n_taps;
    return (int64_t)kWgCtrs * 4 + tiles * a->ksplit * 128 * a->BN * 4;
}

extern "C" int sscg_conv_wgrad(const SscgWgradArgs* a, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->TH * a->TW != 64) return set_error("conv_wgrad: TH*TW must be 64");
    if (a->Kc % a->BN || a->BN % 64) return set_error("conv_wgrad: BN=%d must be a multiple of 64 dividing Kc=%d", a->BN, a->Kc);
    if (a->Co_pad % 64) return set_error("conv_wgrad: Co_pad=%d must be a multiple of 64", a->Co_pad);
    // dy.C < Co_pad and x.C < Kc are allowed: the TMA box zero-fills channels outside the view.
    if (a->split == 3 && (!a->dy_lo || !a->x_lo)) return set_error("conv_wgrad: split=3 needs lo planes");
    if (a->n_taps < 1 || a->n_taps > SSCG_MAX_TAPS) return set_error("conv_wgrad: bad n_taps");
    if (a->ksplit < 1) return set_error("conv_wgrad: ksplit must be >= 1");

    const int rw = a->rw_pitch;
    if (rw != 0) {
        const int G = rw / 16;
        if (rw % 16 || G < 1 || G > 4 || a->x.C != 8 * G || a->x.sW != 8 * G || a->split != 1 || a->stride != 1 || a->TH != 1 ||
            a->TW != 64 || a->BN != 64 * G || a->Kc != 64 * G)
            return set_error("conv_wgrad: pixel-row mode needs a dense view of 8 * G channels (G = 1..4), bf16, stride 1, "
                             "1x64 pixel blocks, BN = Kc = 64 * G");
    }
    CUtensorMap tmDy, tmDyLo, tmX, tmXLo;
    const uint32_t boxDy[4] = {64u, (uint32_t)a->TW, (uint32_t)a->TH, 1u};
    const uint32_t es1[4] = {1u, 1u, 1u, 1u};
    const uint32_t boxX[4] = {64u, (uint32_t)(a->TW * a->stride), (uint32_t)(a->TH * a->stride), 1u};
    const uint32_t esX[4] = {1u, (uint32_t)a->stride, (uint32_t)a->stride, 1u};
    if (int rc = encode_view_4d(&tmDy, a->dy, a->dy.ptr, boxDy, es1)) return rc;
    if (rw != 0) {
        const uint32_t boxR[4] = {8u, (uint32_t)(a->TW + 8), 1u, 1u};
        if (int rc = encode_view_4d(&tmX, a->x, a->x.ptr, boxR, es1, 0)) return rc;
    } else if (int rc = encode_view_4d(&tmX, a->x, a->x.ptr, boxX, esX)) {
        return rc;
    }
    tmDyLo = tmDy; tmXLo = tmX;
    if (a->split == 3) {
        if (int rc = encode_view_4d(&tmDyLo, a->dy, a->dy_lo, boxDy, es1)) return rc;
        if (int rc = encode_view_4d(&tmXLo, a->x, a->x_lo, boxX, esX)) return rc;
    }
    WgradDev d;
    d.N = a->dy.N;
    d.TH = a->TH; d.TW = a->TW;
    d.tiles_h = (a->dy.H + a->TH - 1) / a->TH;
    d.tiles_w = (a->dy.W + a->TW - 1) / a->TW;
    d.stride = a->stride; d.org_h = a->org_h; d.org_w = a->org_w;
    for (int i = 0; i < SSCG_MAX_TAPS; ++i) d.taps[i] = a->taps[i];
    d.Co_pad = a->Co_pad; d.Kc = a->Kc; d.n_ktiles = a->Kc / a->BN;
    d.ksplit = a->ksplit; d.dw = a->dw;
    const int m_tiles = (a->Co_pad + 127) / 128;
    dim3 grid((unsigned)a->ksplit, (unsigned)(m_tiles * d.n_ktiles), (unsigned)a->n_taps);
    d.part = nullptr; d.ctr = nullptr;
    if (a->ksplit > 1) {
        if (a->ws == nullptr) return set_error("conv_wgrad: ksplit > 1 needs the workspace (sscg_conv_wgrad_ws_bytes)");
        if ((long long)grid.y * grid.z * 2 > kWgCtrs) return set_error("conv_wgrad: too many output tiles for the arrival counters");
        d.ctr = reinterpret_cast<unsigned int*>(a->ws);
        d.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->ws) + kWgCtrs * 4);
    }
    if (rw != 0) {
        switch (a->BN) {
            case 64: return launch_wgrad<64, 1, true>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag);
            case 128: return launch_wgrad<128, 1, true>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag);
            case 192: return launch_wgrad<192, 1, true>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag);
            default: return launch_wgrad<256, 1, true>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag);
        }
    }
#define SSCG_WG(BN_)                                                                          \
    case BN_:                                                                                  \
        return a->split == 3 ? launch_wgrad<BN_, 3>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag) \
                             : launch_wgrad<BN_, 1>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag);
    switch (a->BN) {
        SSCG_WG(64)
        SSCG_WG(128)
        SSCG_WG(256)
        case 192:
            if (a->split == 3) return set_error("conv_wgrad: BN=192 is bf16-mode only");
            return launch_wgrad<192, 1>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag);
        case 448:
            if (a->split == 3) return set_error("conv_wgrad: BN=448 is bf16-mode only");
            return launch_wgrad<448, 1>(tmDy, tmDyLo, tmX, tmXLo, d, grid, stream, a->tag);
        default: return set_error("conv_wgrad: unsupported BN=%d", a->BN);
    }
#undef SSCG_WG
}
