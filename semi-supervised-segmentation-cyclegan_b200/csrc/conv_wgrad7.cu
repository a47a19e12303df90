// conv_wgrad7.cu — weight gradient of the 7x7 stride-1 HEAD convolution (64 -> 21 / 19 / 20 / 4 / 3 channels,
// reference arch/generators.py:84-85,89-90) on tcgen05, with the seven horizontal taps as GEMM COLUMNS.
//
//   dW[kh][kw][ci][co] = sum_{n, h, w}  X[n][h + kh][w + kw][ci] * dY[n][h][w][co]          (X carries its halo of 3)
//
// The generic kernel (conv_wgrad.cu, window mode) runs this layer as M = 128 rows of which 21 are output channels and
// streams a 448-wide (kw, ci) window per pixel, i.e. every activation pixel seven times: 631 us per launch at
// 16 x 256 x 256, bound by L2 -> SM traffic (145 B/clk/SM asked for).  Here, with u = w + kw:
//
//   D_kh[ci][(kw, co)] += sum_u  X[n][h + kh][u][ci] * dY[n][h][u - kw][co]
//
// so for one image row h and a block of 64 columns u
//   * A  = X row h + kh            : 64 pixels x 64 channels, MN-major SWIZZLE_128B; two consecutive kh stacked to M = 128
//                                     (two adjacent 8 KB atoms, LBO = 8 KB);
//   * B  = dY row h, ONE segment of 70 pixels x Cy channels (MN-major, SWIZZLE_64B / 32B for Cy = 32 / 16) read as
//          SEVEN OVERLAPPING N-atoms: atom j = 6 - kw starts j pixel rows further down, which the descriptor expresses
//          as a leading-dimension byte offset of ONE row (the swizzle is a pure function of the shared-memory address,
//          so a view shifted by whole rows reads what TMA wrote) — N = 7 * Cy columns from 4.5 KB of shared memory;
//   * one tcgen05.mma (M = 128, N = 224, K = 16 pixels) does the work of 2 x 7 taps.
// A CTA owns four kh (two accumulator pairs, 2 x 224 TMEM columns) and walks h downwards through a ring of X rows:
// row h + kh is loaded once and used by four consecutive K-blocks (ring slots 0..2 are mirrored behind the last slot so
// that the two rows of a pair are always adjacent).  Per K-block 12.5 KB come from L2 for 8 MMAs of 56 clk: 28 B/clk.
// dY sits in a buffer with a ZERO halo of 6 (the layout the N-expanded data gradient needs anyway), which supplies the
// zeros for u - kw outside the row.
//
// Units of work = (kh group, 64-column block, sample, row range); every CTA accumulates its units in TMEM and the
// CTAs of a kh group combine their tiles in CTA order, slice-wise (fixed-order reduction, sscg_ptx.cuh): reproducible.
#include "sscg_common.cuh"

namespace sscg {

constexpr int kW7Ring = 8;                       // X-row ring slots (+3 mirrored)
constexpr int kW7Slot = 64 * 128;                // 64 pixels x 64 channels bf16
constexpr int kW7DySeg = 70;                     // pixels of one dY segment (64 + 6)
constexpr int kW7Ctrs = 64;                      // counters at the head of the workspace

struct Wg7Dev {
    int N, H, W;
    int Cy, row_bytes, NT;      // dY channel pitch (16 / 32), its row bytes, GEMM N = 7 * Cy
    int dy_stage;               // bytes reserved per dY segment (multiple of 512)
    int n_ub, hsplit, units_per_g;
    int ctas_g0, ctas_g1;       // CTAs serving kh 0..3 / kh 4..6
    float* dw;
    float* part;                // [CTA][2 pairs][128 rows][NT]
    unsigned int* ctr;          // [group][arrive, depart]
};

struct W7Unit {
    int u0, n, h0, h1, ksteps;
};
__device__ __forceinline__ W7Unit w7_unit(const Wg7Dev& p, int v) {
    W7Unit t;
    const int per_ub = p.N * p.hsplit;
    const int ub = v / per_ub;
    const int r = v - ub * per_ub;
    t.n = r / p.hsplit;
    const int hs = r - t.n * p.hsplit;
    t.h0 = (int)((long long)hs * p.H / p.hsplit);
    t.h1 = (int)((long long)(hs + 1) * p.H / p.hsplit);
    t.u0 = ub * 64;
    const int valid = min(64, p.W + 6 - t.u0);
    t.ksteps = (valid + 15) >> 4;
    return t;
}

// MN-major operand descriptor: rows (GEMM K = pixels) of row_bytes, 8-row groups contiguous; lbo = byte distance between
// consecutive MN atoms; layout 2 / 4 / 6 = SWIZZLE_128B / 64B / 32B
__device__ __forceinline__ uint64_t w7_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= layout << 61;
    return d;
}

__global__ void __launch_bounds__(192, 1)
conv_wgrad7_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                   const __grid_constant__ Wg7Dev p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xs = smem;                                            // (kW7Ring + 3) slots of 8 KB
    uint8_t* dys = smem + (kW7Ring + 3) * kW7Slot;                 // kW7Ring segments
    uint64_t* bars = reinterpret_cast<uint64_t*>(dys + kW7Ring * p.dy_stage);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kW7Ring;
    uint64_t* tmem_full_bar = bars + 2 * kW7Ring;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kW7Ring + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x < p.ctas_g0 ? 0 : 1;
    const int kh0 = g * 4;
    const int cta_in_g = g == 0 ? blockIdx.x : blockIdx.x - p.ctas_g0;
    const int ctas_g = g == 0 ? p.ctas_g0 : p.ctas_g1;
    const uint32_t tmem_cols = 2 * p.NT <= 256 ? 256u : 512u;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDy);
        for (int s = 0; s < kW7Ring; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(tmem_full_bar), 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();          // the prologue above overlaps the predecessor's tail (launch_k, sscg_common.cuh)
    pdl_launch();
    const bool has_work = cta_in_g < p.units_per_g;

    if (warp == 0) {
        // ================================ TMA producer ==========================================
        if (lane == 0) {
            int T = 0;
            const uint32_t dy_bytes = (uint32_t)(kW7DySeg * p.row_bytes);
            for (int v = cta_in_g; v < p.units_per_g; v += ctas_g) {
                const W7Unit t = w7_unit(p, v);
                const int nsteps = (t.h1 - t.h0) + 3;
                for (int j = 0; j < nsteps; ++j, ++T) {
                    const int slot = T % kW7Ring;
                    mbar_wait(smem_u32(&empty_bar[slot]), (uint32_t)(((T / kW7Ring) & 1) ^ 1), 41);
                    const uint32_t fb = smem_u32(&full_bar[slot]);
                    mbar_arrive_expect_tx(fb, (uint32_t)kW7Slot * (slot < 3 ? 2u : 1u) + (j >= 3 ? dy_bytes : 0u));
                    const int row = t.h0 + kh0 + j;                 // X row (halo coordinates); beyond Hp - 1: zero fill
                    tma_load_4d(smem_u32(xs + slot * kW7Slot), &tmX, fb, 0, t.u0, row, t.n);
                    if (slot < 3) tma_load_4d(smem_u32(xs + (kW7Ring + slot) * kW7Slot), &tmX, fb, 0, t.u0, row, t.n);
                    if (j >= 3) tma_load_4d(smem_u32(dys + slot * p.dy_stage), &tmDy, fb, 0, t.u0, t.h0 + j - 3 + 6, t.n);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ============================================
        if (lane == 0 && has_work) {
            const uint32_t idesc = make_idesc_bf16(128, p.NT, 1, 1);
            const uint64_t blayout = p.row_bytes == 64 ? 4 : 6;
            int T = 0;
            uint32_t accum = 0;
            for (int v = cta_in_g; v < p.units_per_g; v += ctas_g) {
                const W7Unit t = w7_unit(p, v);
                const int nsteps = (t.h1 - t.h0) + 3;
                for (int j = 0; j < nsteps; ++j, ++T) {
                    const int slot = T % kW7Ring;
                    mbar_wait(smem_u32(&full_bar[slot]), (uint32_t)((T / kW7Ring) & 1), 42);
                    tc_fence_after();
                    if (j < 3) continue;                            // the window of four X rows is not complete yet
                    const int s0 = (T - 3) % kW7Ring;               // oldest row of the window: tap row kh0
                    const uint32_t b_addr = smem_u32(dys + slot * p.dy_stage);
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        const uint32_t a_addr = smem_u32(xs + (s0 + 2 * pr) * kW7Slot);
                        // A: two 64-channel atoms (tap rows kh0 + 2 pr, + 1), 8-pixel groups 1 KB apart
                        const uint64_t da = w7_desc(a_addr, kW7Slot, 1024, 2);
                        // B: seven overlapping atoms, one pixel row apart (atom j <-> kw = 6 - j)
                        const uint64_t db = w7_desc(b_addr, (uint32_t)p.row_bytes, (uint32_t)(8 * p.row_bytes), blayout);
                        for (int k = 0; k < t.ksteps; ++k)          // 16 pixels per MMA
                            umma_bf16(tmem_base + pr * p.NT, da + (uint64_t)(128 * k), db + (uint64_t)(p.row_bytes * k),
                                      idesc, (accum | (uint32_t)k) ? 1u : 0u);
                    }
                    accum = 1;
                    umma_commit(smem_u32(&empty_bar[s0]));          // tap row kh0's X row is not needed again
                    if (j == nsteps - 1) {                          // end of the unit: its last three rows as well
                        umma_commit(smem_u32(&empty_bar[(T - 2) % kW7Ring]));
                        umma_commit(smem_u32(&empty_bar[(T - 1) % kW7Ring]));
                        umma_commit(smem_u32(&empty_bar[T % kW7Ring]));
                    }
                }
            }
            umma_commit(smem_u32(tmem_full_bar));
        }
    } else if (has_work) {
        // ================================ epilogue ==============================================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;                 // TMEM lane = accumulator row: (tap row kh0 + 2 pr + m / 64, ci = m % 64)
        const int e = threadIdx.x - 64;
        mbar_wait(smem_u32(tmem_full_bar), 0, 43);
        tc_fence_after();
        const long long tile = 2LL * 128 * p.NT;
        float* mine = p.part + (long long)blockIdx.x * tile;
        for (int pr = 0; pr < 2; ++pr) {
            float* prow = mine + ((long long)pr * 128 + m) * p.NT;
            const uint32_t t_acc = tmem_base + pr * p.NT + (static_cast<uint32_t>(quad * 32) << 16);
            for (int c0 = 0; c0 < p.NT; c0 += 32) {
                uint32_t r[32];
                const int wd = p.NT - c0 >= 32 ? 32 : 16;
                if (wd == 32) tmem_ld_32x32(t_acc + c0, r);
                else tmem_ld_32x16(t_acc + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (4 * q < wd)
                        reinterpret_cast<float4*>(prow + c0)[q] =
                            make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                        __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            }
        }
        tc_fence_before();
        // ---- combine the tiles of this kh group in CTA order, every CTA its own slice ------------------------
        const int nact = min(ctas_g, p.units_per_g);            // CTAs of the group that hold a tile
        unsigned int* ctr_a = p.ctr + 2 * g;
        unsigned int* ctr_d = ctr_a + 1;
        named_bar_sync(1, 128);
        if (e == 0) {
            atom_add_acq_rel_gpu(ctr_a, 1u);
            spin_until_ge(ctr_a, (unsigned int)nact, 44);
        }
        named_bar_sync(2, 128);
        const float* gbase = p.part + (long long)(g == 0 ? 0 : p.ctas_g0) * tile;
        const int total4 = (int)(tile / 4);
        const int per4 = (total4 + nact - 1) / nact;
        const int j0 = cta_in_g * per4, j1 = min(total4, j0 + per4);
        for (int jb = j0 + e; jb < j1; jb += 128) {
            const long long f = 4LL * jb;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = 0; c0 < nact; c0 += 16) {
                float4 vv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    vv[u] = (c0 + u < nact) ? ld_cg_f4(gbase + (long long)(c0 + u) * tile + f) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 16; ++u) { acc.x += vv[u].x; acc.y += vv[u].y; acc.z += vv[u].z; acc.w += vv[u].w; }
            }
            // f -> (pair, row, column) -> (kh, ci, kw, co) -> dW[kh][co][kw * 64 + ci]
            const int pr = (int)(f / (128LL * p.NT));
            const int rem = (int)(f - (long long)pr * 128 * p.NT);
            const int row = rem / p.NT, col = rem - row * p.NT;
            const int kh = kh0 + 2 * pr + (row >> 6), ci = row & 63;
            if (kh > 6) continue;                                // the dummy fourth tap row of group 1
            const int jj = col / p.Cy, co = col - jj * p.Cy;     // 4 consecutive columns share jj (Cy is a multiple of 4)
            const int kw = 6 - jj;
            float* d = p.dw + ((long long)(kh * 64 + co) * 448) + kw * 64 + ci;
            d[0] += acc.x;
            d[448] += acc.y;
            d[896] += acc.z;
            d[1344] += acc.w;
        }
        named_bar_sync(1, 128);
        if (e == 0) {
            const unsigned int old = atom_add_acq_rel_gpu(ctr_d, 1u);
            if (old + 1u == (unsigned int)nact) {
                *ctr_a = 0u;
                *ctr_d = 0u;
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

static int w7_encode(CUtensorMap* tm, const void* ptr, int C, int Wd, int Hd, int N, int box_w, CUtensorMapSwizzle sw) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * Wd, (cuuint64_t)C * 2 * Wd * Hd};
    cuuint32_t b[4] = {(cuuint32_t)C, (cuuint32_t)box_w, 1u, 1u};
    cuuint32_t s[4] = {1, 1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15))
        return set_error("conv_wgrad7: pointers / pitches must be 16-byte aligned");
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, b, s,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error("conv_wgrad7: cuTensorMapEncodeTiled failed: %d (C=%d W=%d H=%d N=%d)", (int)r, C, Wd, Hd, N);
    return 0;
}

static void w7_plan(const SscgWgrad7Args* a, int sms, Wg7Dev& d) {
    d.N = a->N; d.H = a->H; d.W = a->W;
    d.Cy = a->Cy; d.row_bytes = a->Cy * 2; d.NT = 7 * a->Cy;
    d.dy_stage = ((kW7DySeg * d.row_bytes + 511) / 512) * 512;
    d.n_ub = (a->W + 6 + 63) / 64;
    // CTAs: kh 0..3 (two full pairs) and kh 4..6 (+ a dummy row) cost the same; aim at >= ~2 units per CTA
    d.ctas_g0 = sms / 2;
    d.ctas_g1 = sms - d.ctas_g0;
    // row ranges: >= ~8 units per CTA (measured best: 236 us at 16 x 256 x 256 against 276 at 2) so that the static round-robin balances (a unit restart costs three row loads)
    int hsplit = 1;
    const int target = getenv("SSCG_W7_UNITS") ? atoi(getenv("SSCG_W7_UNITS")) : 8;
    while ((long long)d.n_ub * a->N * hsplit < (long long)target * d.ctas_g0 && hsplit * 2 <= a->H / 8) hsplit *= 2;
    d.hsplit = hsplit;
    d.units_per_g = d.n_ub * a->N * hsplit;
}

}  // namespace sscg

using namespace sscg;

static int w7_sms() {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 2) sms = 148;
    return sms;
}

extern "C" int64_t sscg_conv_wgrad7_ws_bytes(const SscgWgrad7Args* a) {
    if (a->Cy != 16 && a->Cy != 32) return -1;
    return (int64_t)kW7Ctrs * 4 + (int64_t)w7_sms() * 2 * 128 * (7 * a->Cy) * 4;
}

extern "C" int sscg_conv_wgrad7(const SscgWgrad7Args* a, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->Cy != 16 && a->Cy != 32) return set_error("conv_wgrad7: Cy=%d must be 16 or 32", a->Cy);
    if (a->N < 1 || a->H < 8 || a->W < 8 || !a->x || !a->dy || !a->dw || !a->ws) return set_error("conv_wgrad7: bad arguments");
    const int sms = w7_sms();
    Wg7Dev d;
    w7_plan(a, sms, d);
    d.dw = a->dw;
    d.ctr = reinterpret_cast<unsigned int*>(a->ws);
    d.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->ws) + kW7Ctrs * 4);
    CUtensorMap tmX, tmDy;
    if (int rc = w7_encode(&tmX, a->x, 64, a->W + 6, a->H + 6, a->N, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = w7_encode(&tmDy, a->dy, a->Cy, a->W + 12, a->H + 12, a->N, kW7DySeg,
                           a->Cy == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    const int smem = 1024 + (kW7Ring + 3) * kW7Slot + kW7Ring * d.dy_stage + 256;
    static int smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_error("conv_wgrad7: cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
        smem_set = smem;
    }
    {
        LaunchScope ls(a->tag, stream);
        launch_k(conv_wgrad7_kernel, sms, 192, smem, stream, tmX, tmDy, d);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("conv_wgrad7 launch: %s", cudaGetErrorString(e));
    return 0;
}
