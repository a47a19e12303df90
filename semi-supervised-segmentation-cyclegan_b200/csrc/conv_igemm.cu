// conv_igemm.cu — implicit-GEMM convolution for sm_100a (persistent, warp-specialised).
//
// Each CTA is resident for the whole launch and walks output tiles (128 pixels x BN channels) with a
// static round-robin schedule.  Roles (192 threads):
//   warp 0      : TMA producer — per K-block one 4-D box of the NHWC activation view (the box *is*
//                 the im2col slice: TH x TW pixels x 64 channels of one filter tap, zero-filled
//                 outside the view, element-strided for stride-2 convolutions) and one 2-D box of
//                 the weight slab, both landing in 128B-swizzled shared memory.  The smem ring runs
//                 across tile boundaries, so the loads of tile i+1 are in flight during tile i.
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue.  Two fp32 accumulators live in
//                 TMEM (2 x BN columns); tcgen05.commit releases smem stages and hands a finished
//                 accumulator to the epilogue while the next tile accumulates into the other one.
//   warps 2..5  : epilogue — tcgen05.ld the accumulator (then immediately give the TMEM buffer back),
//                 optional bias/activation, InstanceNorm partial statistics (warp-shuffle
//                 transpose-reduce per channel, then one order-independent integer accumulation per channel
//                 and tile: binned fixed-point sums, sscg_ptx.cuh), vectorised NHWC store.
//
// Replaces (reference): nn.Conv2d / nn.ConvTranspose2d forward + cuDNN dgrad at
// arch/ops.py:40-57,63,68; arch/generators.py:74-90; arch/discriminators.py:45-58, with
// nn.ReflectionPad2d (ops.py:62,67; generators.py:73,84,89) folded into the operand view and the
// statistics of nn.InstanceNorm2d (ops.py:11) fused into the epilogue.
#include "sscg_common.cuh"

namespace sscg {

struct ConvDev {
    int N, Ho, Wo;
    int n_phases;
    int phase_start[5];
    SscgTap taps[SSCG_MAX_TAPS];
    int stride, org_h, org_w;
    int kcb;        // 64-wide K blocks per tap
    int Co_pad;
    int TH, TW, tw_shift;
    int tiles_h, tiles_w;
    int shift_brow_step;   // row-shift mode: weight slab step between consecutive kw taps (+1 / -1)
    int shift_base_mode;   // row-shift mode: 1 = descriptor base_offset carries the row phase, 2 = base_offset 0
    int tiles_x;    // tiles_h * tiles_w * N
    int n_ntiles;   // Co_pad / BN
    int total_tiles;
    int flat_pitch, flat_hw;   // flattened tiling (see SscgConvArgs): input row pitch / positions per sample; 0 = off
    int tail_from;  // >= 0: schedule entries from this index on are HALF-N tiles (two per output tile): the last,
                    // partial wave of a persistent grid then costs ~0.6 instead of 1.0 tile times (BN = 256 only)
    void* y;
    int y_fp32;
    long long y_sN, y_sH, y_sW;
    int y_oh, y_ow;
    const float* bias;
    int act;
    float slope;
    unsigned long long* stats;   // [N][Co_pad][2][kDetWords] binned accumulators of (sum, sum of squares)
};

#ifndef SSCG_IGEMM_BN128_KB
#define SSCG_IGEMM_BN128_KB 200
#endif
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;   // 128 pixels x 64 bf16

// SKW > 0 selects the row-shift mode (1 x 128 pixel tiles, stride 1): one TMA box of 128 + SKW - 1
// pixels per filter row and channel block is shared by the SKW horizontal taps, which read it through
// descriptors whose start address is shifted by one 128-byte pixel row per tap.
//
// RW > 0 selects the PIXEL-ROW mode for stems (few input channels, 1 x 128 pixel tiles, stride 1; RW = 16 = bytes per pixel of
// ONE 8-channel group of the input buffer; inputs with 16 / 24 / 32 channels run one K block per group, the TMA box
// picking the group's 8 channels out of every pixel): the K axis of one filter row is the run of kw consecutive pixels x Cp channels, which IS the image
// row.  One TMA box of 128 + 8 pixels lands densely (16 bytes per pixel for Cp = 8) and the A descriptor addresses it
// as overlapping rows: M row m starts at pixel m (16-byte row pitch inside a core matrix, SBO = 128 B per 8 pixels),
// the next 16-byte K chunk is the next pixel (LBO = 16 B).  The earlier row-WINDOW tensor map made TMA expand every
// pixel's 128-byte window (8x the unique bytes: 1.5 GB of L2 -> SM traffic for a 17 MB input, the kernel's bound).
template <int BN, int SPLIT, int SKW = 0, int RW = 0>
struct IgemmCfg {
    static constexpr int kPlanes = (SPLIT == 3) ? 2 : 1;
    static constexpr int kBBytes = BN * 128;
    static constexpr int kARows = RW > 0 ? kTileM + 8 : (SKW > 0 ? kTileM + SKW - 1 : kTileM);
    static constexpr int kARowBytes = RW > 0 ? RW : 128;
    static constexpr int kAStage = (SKW > 0 || RW > 0) ? ((kARows * kARowBytes + 1023) / 1024) * 1024 : kABytes;
    static constexpr int kBStage = SKW > 0 ? SKW * kBBytes : kBBytes;
    static constexpr int kStageBytes = kPlanes * (kAStage + kBStage);
    static constexpr int kTxBytes = kPlanes * (kARows * kARowBytes + kBStage);
    // wide tiles: as deep a ring as fits one CTA per SM; narrow tiles: 4 stages so that 2+ CTAs fit
    // SSCG_IGEMM_BN128_KB: ring budget of the BN = 128 tile (200: one CTA per SM with 6 stages; 100: 3 stages, two CTAs)
    static constexpr int kMaxStages = (RW > 0 ? 100 * 1024 : (BN > 128 || SKW > 0) ? 200 * 1024
                                       : (BN == 128 ? SSCG_IGEMM_BN128_KB * 1024 : 100 * 1024)) / kStageBytes;
    static constexpr int kStages = kMaxStages > 6 ? 6 : (kMaxStages < 2 ? 2 : kMaxStages);
    static constexpr int kAccCols = BN < 32 ? 32 : BN;          // columns per accumulator
    static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;
    static constexpr int kStatBytes = 4 * BN * 2 * 4;
    static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kStatBytes + 256;
};

struct TileInfo {
    int n, i0, j0, n0, ph, pw, Hph, Wph, tap0, nkb;
    int bn;         // columns of this tile: BN, or BN / 2 for a tail tile
    bool valid;
};

__device__ __forceinline__ TileInfo decode_tile(const ConvDev& p, int tile, int BN) {
    TileInfo t;
    int half = -1;
    if (p.tail_from >= 0 && tile >= p.tail_from) {
        const int h = tile - p.tail_from;
        tile = p.tail_from + (h >> 1);
        half = h & 1;
    }
    int x = tile % p.tiles_x;
    int rest = tile / p.tiles_x;
    const int y = rest % p.n_ntiles;
    const int z = rest / p.n_ntiles;
    const int os = (p.n_phases == 4) ? 2 : 1;
    t.ph = (p.n_phases == 4) ? (z >> 1) : 0;
    t.pw = (p.n_phases == 4) ? (z & 1) : 0;
    t.Hph = (p.Ho - t.ph + os - 1) / os;
    t.Wph = (p.Wo - t.pw + os - 1) / os;
    const int tj = x % p.tiles_w; x /= p.tiles_w;
    const int ti = x % p.tiles_h; x /= p.tiles_h;
    t.n = x;
    t.i0 = ti * p.TH;
    t.j0 = tj * p.TW;
    t.n0 = y * BN;
    t.bn = BN;
    if (half >= 0) {
        t.bn = BN >> 1;
        t.n0 += half * t.bn;
    }
    t.tap0 = p.phase_start[z];
    t.nkb = (p.phase_start[z + 1] - t.tap0) * p.kcb;
    t.valid = p.flat_pitch > 0 ? true : ((t.i0 < t.Hph) && (t.j0 < t.Wph));
    return t;
}

// Schedule of one CTA: a CONTIGUOUS range of the full-tile entries, then its share of the half-N tail entries (which
// keep the strided order: they are the balancing last wave).  Consecutive tiles mostly belong to one sample, so the CTA
// keeps the plane sums of its tiles in registers and adds them to the global accumulators once per sample change: with
// one flush per tile the 64-bit reductions cost 6 of the 67 us of a residual-block convolution (524 K per launch) and
// 45 of the 164 us of the 64-channel stem, whose 512 tiles per sample all hit the same 256 words.
struct TileWalk {
    int cur, end, tail, total, step;
    __device__ __forceinline__ explicit TileWalk(const ConvDev& p) {
        // four-phase launches (transposed conv / stride-2 gathers) keep the strided order for ALL entries: their
        // phases have 1, 2, 2 and 4 taps, and a contiguous range would fall into one phase (72 -> 90 us)
        const int F = p.n_phases == 4 ? 0 : (p.tail_from >= 0 ? p.tail_from : p.total_tiles);
        const int G = (int)gridDim.x, b = (int)blockIdx.x;
        const int q = F / G, r = F - q * G;
        cur = b * q + (b < r ? b : r);
        end = cur + q + (b < r ? 1 : 0);
        tail = F + b;
        total = p.total_tiles;
        step = G;
    }
    __device__ __forceinline__ int next() {      // -1: done
        if (cur < end) return cur++;
        if (tail < total) {
            const int t = tail;
            tail += step;
            return t;
        }
        return -1;
    }
};

template <int BN, int SPLIT, int SKW, int RW>
__global__ void __launch_bounds__(192, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                  const __grid_constant__ ConvDev p) {
    using Cfg = IgemmCfg<BN, SPLIT, SKW, RW>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kPlanes = Cfg::kPlanes;
    static_assert(SKW == 0 || SPLIT == 1, "row-shift mode is bf16-mode only");
    static_assert(RW == 0 || (RW == 16 && SPLIT == 1 && SKW == 0), "pixel-row mode: 16 bytes per pixel, bf16 mode");

    // ---- shared memory carve-up ----------------------------------------------------------------
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* s_stat = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kStatBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full_bar = bars + 2 * kStages;        // [2]
    uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (SPLIT == 3) { tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmBlo); }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tmem_full_bar[s]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[s]), 128);   // every epilogue thread arrives
        }
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

    if (warp == 0) {
        // ================================ TMA producer ==========================================
        if (lane == 0) {
            int stage = 0; uint32_t par = 0;
            TileWalk tw(p);
            for (int tile = tw.next(); tile >= 0; tile = tw.next()) {
                const TileInfo t = decode_tile(p, tile, BN);
                if (!t.valid) continue;
                int tp = 0, cb = 0;
                for (int kb = 0; kb < t.nkb; ++kb) {
                    const SscgTap tap = p.taps[t.tap0 + tp];
                    mbar_wait(smem_u32(&empty_bar[stage]), par ^ 1, 1);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    const bool half_tile = (SPLIT == 1 && SKW == 0 && RW == 0) && t.bn != BN;
                    mbar_arrive_expect_tx(fb, half_tile ? Cfg::kTxBytes - Cfg::kBBytes / 2 : Cfg::kTxBytes);
                    uint8_t* st = smem + stage * Cfg::kStageBytes;
                    int cw = t.j0 * p.stride + tap.dw + p.org_w;
                    int ch = t.i0 * p.stride + tap.dh + p.org_h;
                    int cn = t.n;
                    if (p.flat_pitch > 0) {      // 1-D pixel view: tile start + tap offset in flattened positions
                        cw = t.n * p.flat_hw + t.j0 + tap.dh * p.flat_pitch + tap.dw;
                        ch = 0;
                        cn = 0;
                    }
                    const int brow = tap.brow * p.Co_pad + t.n0;
                    tma_load_4d(smem_u32(st), &tmA, fb, RW > 0 ? cb * 8 : cb * 64, cw, ch, cn);   // RW: channel group cb
                    if (SKW > 0) {   // one weight box per horizontal tap of this filter row
#pragma unroll
                        for (int j = 0; j < (SKW > 0 ? SKW : 1); ++j)
                            tma_load_2d(smem_u32(st + Cfg::kAStage + j * Cfg::kBBytes), &tmB, fb, cb * 64,
                                        (tap.brow + j * p.shift_brow_step) * p.Co_pad + t.n0);
                    } else if (RW > 0) {
                        tma_load_2d(smem_u32(st + Cfg::kAStage), &tmB, fb, cb * 64, brow);
                    } else {
                        // half tiles read their 128 weight rows through the half-height box (passed in the tmBlo slot)
                        tma_load_2d(smem_u32(st + kPlanes * kABytes), half_tile ? &tmBlo : &tmB, fb, cb * 64, brow);
                    }
                    if (SPLIT == 3) {
                        tma_load_4d(smem_u32(st + kABytes), &tmAlo, fb, cb * 64, cw, ch, cn);
                        tma_load_2d(smem_u32(st + kPlanes * kABytes + Cfg::kBBytes), &tmBlo, fb, cb * 64, brow);
                    }
                    if (++stage == kStages) { stage = 0; par ^= 1; }
                    if (++cb == p.kcb) { cb = 0; ++tp; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ============================================
        if (lane == 0) {
            constexpr uint32_t idesc_full = make_idesc_bf16(kTileM, BN < 16 ? 16 : BN, 0, 0);
            constexpr uint32_t idesc_half = make_idesc_bf16(kTileM, BN < 32 ? 16 : BN / 2, 0, 0);
            int stage = 0; uint32_t par = 0;
            uint32_t it = 0;
            TileWalk tw(p);
            for (int tile = tw.next(); tile >= 0; tile = tw.next()) {
                const TileInfo t = decode_tile(p, tile, BN);
                if (!t.valid) continue;
                const uint32_t acc = it & 1, acc_par = (it >> 1) & 1;
                ++it;
                mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_par ^ 1, 7);   // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
                const uint32_t idesc = (t.bn != BN) ? idesc_half : idesc_full;
                uint32_t accum = 0;
                for (int kb = 0; kb < t.nkb; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), par, 2);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    if (SKW > 0) {
                        const uint32_t sbw = sa + Cfg::kAStage;
#pragma unroll
                        for (int j = 0; j < (SKW > 0 ? SKW : 1); ++j) {
                            // tap j reads pixel rows j .. j+127 of the shared row box
                            const uint64_t daj = make_smem_desc_sw128(sa + j * 128, 0, 1024,
                                                                      p.shift_base_mode == 1 ? (uint32_t)j : 0u);
                            const uint64_t dbj = make_smem_desc_sw128(sbw + j * Cfg::kBBytes, 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_bf16(d_tmem, daj + 2 * k, dbj + 2 * k, idesc, accum);
                                accum = 1;
                            }
                        }
                        umma_commit(smem_u32(&empty_bar[stage]));
                        if (++stage == kStages) { stage = 0; par ^= 1; }
                        continue;
                    }
                    const uint32_t sb = sa + (RW > 0 ? Cfg::kAStage : kPlanes * kABytes);
                    // pixel-row mode: dense 16-byte pixels, no swizzle; K chunk = next pixel (LBO 16), 8 rows = 8 pixels (SBO 128)
                    const uint64_t da = RW > 0 ? make_smem_desc(sa, 16, 128, 0) : make_smem_desc_sw128(sa, 0, 1024);
                    const uint64_t db = make_smem_desc_sw128(sb, 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // 4 x (K = 16) per 64-wide K block; +32 B per step
                        umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                        accum = 1;
                        if (SPLIT == 3) {
                            const uint64_t dalo = make_smem_desc_sw128(sa + kABytes, 0, 1024);
                            const uint64_t dblo = make_smem_desc_sw128(sb + Cfg::kBBytes, 0, 1024);
                            umma_bf16(d_tmem, dalo + 2 * k, db + 2 * k, idesc, 1);
                            umma_bf16(d_tmem, da + 2 * k, dblo + 2 * k, idesc, 1);
                        }
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));   // stage reusable once these MMAs retire
                    if (++stage == kStages) { stage = 0; par ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar[acc]));     // accumulator complete
            }
        }
    } else {
        // ================================ epilogue ==============================================
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int m = quad * 32 + lane;            // accumulator row = pixel within the tile
        const int pi = m >> p.tw_shift, pj = m & (p.TW - 1);
        const int os = (p.n_phases == 4) ? 2 : 1;
        constexpr int kChunks = (BN + 31) / 32;
        constexpr int kCols = BN >= 32 ? 32 : BN;
        uint32_t it = 0;
        // running plane sums of this thread's columns (threads 0..127 of the epilogue own columns e, e + 128)
        float run1[2] = {0.f, 0.f}, run2[2] = {0.f, 0.f};
        int run_n = -1, run_n0 = 0, run_bn = 0;
        auto flush_stats = [&]() {
            if (run_n < 0) return;
            const int e = threadIdx.x - 64;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int col = e + 128 * k;
                if (col < run_bn) {
                    // binned integer accumulators (sscg_ptx.cuh): order-independent, hence reproducible, plane sums
                    unsigned long long* dst = p.stats + ((long long)run_n * p.Co_pad + run_n0 + col) * (2 * kDetWords);
                    det_red_add(dst, run1[k]);
                    det_red_add(dst + kDetWords, run2[k]);
                }
                run1[k] = run2[k] = 0.f;
            }
        };
        TileWalk tw(p);
        for (int tile = tw.next(); tile >= 0; tile = tw.next()) {
            const TileInfo t = decode_tile(p, tile, BN);
            if (!t.valid) continue;
            const uint32_t acc = it & 1, acc_par = (it >> 1) & 1;
            ++it;
            int i = t.i0 + pi, j = t.j0 + pj;
            if (p.flat_pitch > 0) {            // position f of the flattened sample -> (row, col); halo columns are dropped
                const int f = t.j0 + m;
                i = f / p.flat_pitch;
                j = f - i * p.flat_pitch;
            }
            const bool valid = (i < t.Hph) && (j < t.Wph);
            const int ho = i * os + t.ph, wo = j * os + t.pw;
            const long long yoff = (long long)t.n * p.y_sN + (long long)(ho + p.y_oh) * p.y_sH +
                                   (long long)(wo + p.y_ow) * p.y_sW + t.n0;
            mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_par, 3);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + acc * Cfg::kAccCols + (static_cast<uint32_t>(quad * 32) << 16);
            const int nch = (t.bn != BN) ? kChunks / 2 : kChunks;
#pragma unroll 1
            for (int c = 0; c < nch; ++c) {
                uint32_t r[32];
                if (BN >= 32) {
                    tmem_ld_32x32(t_acc + c * 32, r);
                } else {
                    tmem_ld_32x16(t_acc + c * 32, r);
#pragma unroll
                    for (int q = 16; q < 32; ++q) r[q] = 0;
                }
                tmem_ld_wait();
                if (c == nch - 1) {                // accumulator fully read: hand the TMEM buffer back
                    tc_fence_before();
                    mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
                }
                float v[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
                if (p.bias != nullptr) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] += __ldg(p.bias + t.n0 + c * 32 + q);
                }
                if (p.act == SSCG_ACT_RELU) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.f);
                } else if (p.act == SSCG_ACT_LRELU) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * p.slope;
                } else if (p.act == SSCG_ACT_TANH) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = tanhf(v[q]);
                }
                if (p.y_fp32) {
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + yoff + c * 32);
#pragma unroll
                        for (int q = 0; q < kCols / 4; ++q)
                            dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                } else {
                    uint32_t pk[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) pk[q] = pack_bf16x2(v[2 * q], v[2 * q + 1]);
                    if (valid) {
                        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yoff + c * 32);
#pragma unroll
                        for (int q = 0; q < kCols / 8; ++q)
                            dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    }
                    if (p.stats != nullptr) {   // statistics of the values as stored
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            v[2 * q] = __uint_as_float(pk[q] << 16);
                            v[2 * q + 1] = __uint_as_float(pk[q] & 0xffff0000u);
                        }
                    }
                }
                if (p.stats != nullptr) {
                    float s1[32], s2[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float x = valid ? v[q] : 0.f;
                        s1[q] = x;
                        s2[q] = x * x;
                    }
                    // transpose-reduce over the 32 lanes (pixels): lane L ends with column L's sum
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int q = 0; q < off; ++q) {
                            const float send1 = up ? s1[q] : s1[q + off];
                            const float keep1 = up ? s1[q + off] : s1[q];
                            s1[q] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
                            const float send2 = up ? s2[q] : s2[q + off];
                            const float keep2 = up ? s2[q + off] : s2[q];
                            s2[q] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
                        }
                    }
                    const int col = c * 32 + lane;
                    if (col < t.bn) {
                        s_stat[(quad * BN + col) * 2 + 0] = s1[0];
                        s_stat[(quad * BN + col) * 2 + 1] = s2[0];
                    }
                }
            }
            if (p.stats != nullptr) {
                named_bar_sync(1, 128);               // all four quadrants wrote their partials
                if (t.n != run_n || t.n0 != run_n0 || t.bn != run_bn) {
                    flush_stats();
                    run_n = t.n; run_n0 = t.n0; run_bn = t.bn;
                }
                const int e = threadIdx.x - 64;       // 0..127
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int col = e + 128 * k;
                    if (col < t.bn) {
                        float a = 0.f, b = 0.f;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            a += s_stat[(q * BN + col) * 2 + 0];
                            b += s_stat[(q * BN + col) * 2 + 1];
                        }
                        run1[k] += a;
                        run2[k] += b;
                    }
                }
                named_bar_sync(2, 128);               // s_stat may be overwritten by the next tile
            }
        }
        if (p.stats != nullptr) flush_stats();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN, int SPLIT, int SKW = 0, int RW = 0>
static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmAlo, const CUtensorMap& tmB,
                        const CUtensorMap& tmBlo, const ConvDev& d, cudaStream_t stream, int tag) {
    using Cfg = IgemmCfg<BN, SPLIT, SKW, RW>;
    static int occ = 0;
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN, SPLIT, SKW, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmemBytes);
        if (e != cudaSuccess) return set_error("conv_igemm: cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes,
                                               cudaGetErrorString(e));
        // without an explicit carve-out the driver sizes shared memory for ONE block of this kernel and the
        // occupancy query answers 1 even when two blocks would fit the 228 KB
        cudaFuncSetAttribute(conv_igemm_kernel<BN, SPLIT, SKW, RW>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, conv_igemm_kernel<BN, SPLIT, SKW, RW>, 192, Cfg::kSmemBytes);
        const int o_api = o;
        const int e_api = (int)e;
        if (e != cudaSuccess || o < 1) o = 1;
        // The occupancy query answers 1 for these kernels (driver 580) even where two blocks fit by shared
        // memory, registers and TMEM columns (ncu: launch__occupancy_limit_* = 2; a second resident CTA measurably
        // overlaps one tile's epilogue with the other's main loop: +20..40 % on the narrow small-K layers).
        // Size the grid by shared memory; registers (<= 168 x 192 threads) and TMEM are checked here.
        {
            const int by_smem = (int)(233472 / (Cfg::kSmemBytes + 1024));
            const char* ov = getenv("SSCG_IGEMM_OCC");
            if (by_smem > o && !(ov && atoi(ov) == 1)) o = by_smem > 2 ? 2 : by_smem;
        }
        const int tmem_limit = 512 / Cfg::kTmemCols;     // resident CTAs must all fit their TMEM columns
        if (o > tmem_limit) o = tmem_limit;
        if (o > 4) o = 4;
        occ = o < 1 ? 1 : o;
        if (getenv("SSCG_DEBUG"))
            fprintf(stderr, "[sscg] conv_igemm<%d,%d,%d>: smem %d B, stages %d, tmem cols %d, occupancy %d (api %d, err %d)\n", BN, SPLIT,
                    SKW, Cfg::kSmemBytes, Cfg::kStages, Cfg::kTmemCols, occ, o_api, e_api);
    }
    int grid = sm_count() * occ;
    if (grid > d.total_tiles) grid = d.total_tiles;
    {
        LaunchScope ls(tag, stream);
        launch_k(conv_igemm_kernel<BN, SPLIT, SKW, RW>, grid, 192, Cfg::kSmemBytes, stream, tmA, tmAlo, tmB, tmBlo, d);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("conv_igemm<%d,%d> launch: %s", BN, SPLIT, cudaGetErrorString(e));
    return 0;
}

}  // namespace sscg

using namespace sscg;

extern "C" int sscg_conv_igemm(const SscgConvArgs* a, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->TH * a->TW != 128 || (a->TW & (a->TW - 1))) return set_error("conv_igemm: TH*TW must be 128, TW a power of two");
    if (a->Kc % 64) return set_error("conv_igemm: Kc=%d must be a multiple of 64", a->Kc);   // x.C < Kc: zero-filled
    if (a->stride != 1 && a->stride != 2) return set_error("conv_igemm: stride must be 1 or 2");
    if (a->n_phases != 1 && a->n_phases != 4) return set_error("conv_igemm: n_phases must be 1 or 4");
    if (a->Co_pad % a->BN) return set_error("conv_igemm: Co_pad=%d not a multiple of BN=%d", a->Co_pad, a->BN);
    if (a->split != 1 && a->split != 3) return set_error("conv_igemm: split must be 1 or 3");
    if (a->split == 3 && (!a->x_lo || !a->w_lo)) return set_error("conv_igemm: split=3 needs lo planes");
    if (a->phase_start[a->n_phases] > SSCG_MAX_TAPS) return set_error("conv_igemm: too many taps");
    if (a->stride * a->TW > 256 || a->stride * a->TH > 256) return set_error("conv_igemm: box too large");

    const int skw = a->shift_kw;
    if (skw != 0) {
        if (skw != 7 || a->split != 1 || a->stride != 1 || a->n_phases != 1 || a->TH != 1 || a->TW != 128 ||
            (a->BN != 16 && a->BN != 32 && a->BN != 64))
            return set_error("conv_igemm: row-shift mode needs kw=7, bf16, stride 1, 1x128 tiles, BN 16/32/64");
    }
    const int rw = a->rw_pitch;
    if (rw != 0) {
        const int G = rw / 16;
        if (rw % 16 || G < 1 || G > 4 || a->x.C != 8 * G || a->x.sW != 8 * G || a->split != 1 || a->stride != 1 ||
            a->n_phases != 1 || a->TH != 1 || a->TW != 128 || a->BN != 64 || a->Kc != 64 * G || skw != 0 || a->flat_pitch > 0)
            return set_error("conv_igemm: pixel-row mode needs a dense view of 8 * G channels (G = 1..4), bf16, stride 1, "
                             "1x128 tiles, BN 64, Kc = 64 * G");
    }
    CUtensorMap tmA, tmAlo, tmB, tmBlo;
    const uint32_t boxA[4] = {rw ? 8u : 64u, (uint32_t)(rw ? a->TW + 8 : (skw ? a->TW + skw - 1 : a->TW * a->stride)),
                              (uint32_t)(a->TH * a->stride), 1u};
    const uint32_t esA[4] = {1u, (uint32_t)a->stride, (uint32_t)a->stride, 1u};
    if (int rc = encode_view_4d(&tmA, a->x, a->x.ptr, boxA, esA, rw ? 0 : 128)) return rc;
    tmAlo = tmA;
    if (a->split == 3)
        if (int rc = encode_view_4d(&tmAlo, a->x, a->x_lo, boxA, esA)) return rc;
    if (int rc = encode_2d(&tmB, a->w, a->Kc, a->w_rows, 64, a->BN)) return rc;
    tmBlo = tmB;
    if (a->split == 3)
        if (int rc = encode_2d(&tmBlo, a->w_lo, a->Kc, a->w_rows, 64, a->BN)) return rc;

    ConvDev d;
    d.N = a->x.N; d.Ho = a->Ho; d.Wo = a->Wo;
    d.n_phases = a->n_phases;
    for (int i = 0; i < 5; ++i) d.phase_start[i] = a->phase_start[i];
    for (int i = 0; i < SSCG_MAX_TAPS; ++i) d.taps[i] = a->taps[i];
    d.stride = a->stride; d.org_h = a->org_h; d.org_w = a->org_w;
    d.kcb = a->Kc / 64;
    d.Co_pad = a->Co_pad;
    d.TH = a->TH; d.TW = a->TW;
    d.tw_shift = 0; while ((1 << d.tw_shift) < a->TW) ++d.tw_shift;
    const int os = a->n_phases == 4 ? 2 : 1;
    const int Hph = (a->Ho + os - 1) / os, Wph = (a->Wo + os - 1) / os;
    d.tiles_h = (Hph + a->TH - 1) / a->TH;
    d.tiles_w = (Wph + a->TW - 1) / a->TW;
    d.flat_pitch = a->flat_pitch;
    d.flat_hw = a->flat_hw;
    if (a->flat_pitch > 0) {
        if (a->stride != 1 || a->n_phases != 1 || skw != 0 || a->TH != 1 || a->TW != 128 || a->flat_n < 1 ||
            a->flat_pitch < a->Wo || a->stats != nullptr)
            return set_error("conv_igemm: flattened tiling needs stride 1, one phase, 1x128 tiles, pitch >= Wo, no statistics");
        d.N = a->flat_n;
        d.tiles_h = 1;
        d.tiles_w = (a->Ho * a->flat_pitch + 127) / 128;     // positions of rows 0 .. Ho-1, halo columns included
    }
    d.shift_brow_step = a->shift_brow_step;
    d.shift_base_mode = a->shift_base_mode;
    d.tiles_x = d.tiles_h * d.tiles_w * d.N;
    d.n_ntiles = a->Co_pad / a->BN;
    d.total_tiles = d.tiles_x * d.n_ntiles * a->n_phases;
    d.tail_from = -1;
    if (a->BN == 256 && a->split == 1 && skw == 0 && !getenv("SSCG_NO_TAIL_SPLIT")) {
        // persistent grid of one CTA per SM: if the last wave is at most half full, run it as half-N tiles
        const int G = sm_count(), T = d.total_tiles, rem = T % G;
        if (T > G && rem > 0 && 2 * rem <= G) {
            d.tail_from = T - rem;
            d.total_tiles = T + rem;
            if (int rc = encode_2d(&tmBlo, a->w, a->Kc, a->w_rows, 64, a->BN / 2)) return rc;
        }
    }
    d.y = a->y; d.y_fp32 = a->y_fp32;
    d.y_sN = a->y_sN; d.y_sH = a->y_sH; d.y_sW = a->y_sW; d.y_oh = a->y_oh; d.y_ow = a->y_ow;
    d.bias = a->bias; d.act = a->act; d.slope = a->slope; d.stats = reinterpret_cast<unsigned long long*>(a->stats);
    if (d.total_tiles <= 0) return 0;
    if (rw != 0) return launch_igemm<64, 1, 0, 16>(tmA, tmAlo, tmB, tmBlo, d, stream, a->tag);
    if (skw == 7) {
        if (a->BN == 16) return launch_igemm<16, 1, 7>(tmA, tmAlo, tmB, tmBlo, d, stream, a->tag);
        if (a->BN == 64) return launch_igemm<64, 1, 7>(tmA, tmAlo, tmB, tmBlo, d, stream, a->tag);
        return launch_igemm<32, 1, 7>(tmA, tmAlo, tmB, tmBlo, d, stream, a->tag);
    }

#define SSCG_DISPATCH(BN_)                                                                  \
    case BN_:                                                                                \
        return a->split == 3 ? launch_igemm<BN_, 3>(tmA, tmAlo, tmB, tmBlo, d, stream, a->tag) \
                             : launch_igemm<BN_, 1>(tmA, tmAlo, tmB, tmBlo, d, stream, a->tag);
    switch (a->BN) {
        SSCG_DISPATCH(16)
        SSCG_DISPATCH(32)
        SSCG_DISPATCH(64)
        SSCG_DISPATCH(128)
        SSCG_DISPATCH(256)
        default: return set_error("conv_igemm: unsupported BN=%d", a->BN);
    }
#undef SSCG_DISPATCH
}
