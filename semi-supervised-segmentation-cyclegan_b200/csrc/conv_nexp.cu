// conv_nexp.cu — 7x7 stride-1 convolutions with a NARROW output (the generator head, 64 -> 21 / 3 channels,
// and its data gradient, 21 / 3 -> 64 channels) as an "N-expanded" implicit GEMM on tcgen05.
//
// The regular kernel (conv_igemm.cu) runs these layers as M = 128 pixels x N = Cout (16..64) x K = 49 * 64:
// 196 MMAs per tile whose cost is set by the shared-memory read of the 128 x 16 A slab (~60 cycles each,
// whatever N is), i.e. the tensor pipe sits at 10-40 %.  Here the seven HORIZONTAL taps move from K to N:
//
//     P[p][(kw, co)] = sum_{kh, c} X[row(p) + kh - 3][col(p)][c] * W[kh][kw][c][co]        (GEMM, K = 7 * 64)
//     Y[o][co]       = sum_{kw}    P[o + kw - 3][(kw, co)]                                   (epilogue)
//
// so a tile needs 7 * ksteps MMAs of N = 7 * CoW columns (up to 224) instead of 196 narrow ones, and the
// operand read per useful MAC drops ~6x.  Pixels are addressed in the FLATTENED padded image (p = row * Wp + col):
// a shift by kw - 3 is a shift of the flattened index, wrong only across row ends, which are halo columns whose
// outputs are never stored.  A tile is 128 consecutive flattened pixels and yields 122 outputs; the shift-add
// runs through a 128 x (N + 1) fp32 staging buffer in shared memory (conflict-free for both the row-wise store
// and the diagonal read).
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer (two TMEM accumulators), and TWO epilogue groups of four warps
// (warps 2..5 and 6..9), one per accumulator, each with a staging buffer of its own: the epilogue of a tile — 115 KB
// out of TMEM (64 B/clk), 230 KB through shared memory for the shift-add — takes ~5 k cycles against a 0.8 k main loop,
// so a single group set the pace of the whole kernel (366 us for the head's data gradient at 16 x 256 x 256); two
// groups overlap the TMEM phase of one tile with the shared-memory phase of the other.
// Replaces (reference): the 7x7 nn.Conv2d head of arch/generators.py:84-85,89-90 (forward and cuDNN dgrad).
#include "sscg_common.cuh"

namespace sscg {

struct NexpDev {
    int N;              // samples
    int Hp, Wp;         // padded input extents (flattened pixel space of one sample: Hp * Wp)
    int tiles_per_sample, n_ntiles, total_tiles;
    int NT;             // GEMM N of one tile = round_up(7 * CoW, 16)
    int CoW;            // output columns per horizontal tap in one N tile
    int ksteps;         // 16-channel K steps per vertical tap (1, 2 or 4)
    int row_bytes;      // 32 * ksteps: operand rows are exactly as wide as the channels that enter the contraction,
                        // with the matching swizzle (128B / 64B / 32B) for TMA and the UMMA descriptors
    int stages;
    int wstat;          // weight-stationary: the seven vertical-tap slabs of this CTA's N tile stay in shared memory for
                        // the whole launch (they were 2/3 of the L2 -> SM traffic of a tile); needs gridDim.x % n_ntiles == 0
    int halves;         // 1: the staging buffer holds all NT columns; p > 1: the shift-add runs in p passes of CoW / p
                        // channels (a multiple of 8) through a buffer of 7 * CoW / p columns (room for resident weights)
    int groups;         // epilogue groups (2 when two staging buffers fit, else 1)
    int s_bytes;        // bytes of one staging buffer
    void* y;
    int y_fp32;
    long long y_sN, y_sH, y_sW;
    int y_c0_step;      // output channel offset per N tile (= CoW)
    int c_store;        // channels written per pixel (<= CoW, multiple of 8)
    const float* bias;
    int act;
};

constexpr int kNxTileM = 128;
constexpr int kNxOut = 122;               // outputs per tile (128 - 6 halo pixels)

// K-major operand descriptor for rows of 128 / 64 / 32 bytes (SWIZZLE_128B / 64B / 32B): 8-row atoms, SBO = 8 rows
__device__ __forceinline__ uint64_t nexp_desc(uint32_t saddr, int row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(((8u * row_bytes) >> 4) & 0x3fff) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= layout << 61;
    return d;
}

// explicit shared-space accesses: the staging buffer is carved out of dynamic shared memory through integer
// arithmetic, which makes the compiler fall back to generic LD / ST (address translation, long scoreboard)
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

__global__ void __launch_bounds__(320, 1)
conv_nexp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ NexpDev p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = kNxTileM * p.row_bytes;
    const int b_bytes = p.NT * p.row_bytes;
    const int a_al = (a_bytes + 1023) & ~1023, b_al = (b_bytes + 1023) & ~1023;
    const int stage_bytes = p.wstat ? a_al : a_al + b_al;
    const int b_off = a_al;
    uint8_t* wres = smem + p.stages * stage_bytes;                       // resident weight slabs (wstat): 7 x b_al
    float* S = reinterpret_cast<float*>(wres + (p.wstat ? 7 * b_al : 0));
    const int ch_half = p.CoW / p.halves;                                // channels per shift-add pass
    const int s_cols = p.halves == 1 ? p.NT : 7 * ch_half;               // staged columns per pass
    const int s_pitch = s_cols + 1;                     // odd: conflict-free row-wise stores and diagonal reads
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(S) + p.groups * p.s_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 4;
    uint64_t* tmem_full_bar = bars + 8;        // [2]
    uint64_t* tmem_empty_bar = bars + 10;      // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);
    uint64_t* w_bar = bars + 13;               // resident weights have landed
    float* bias_sm = reinterpret_cast<float*>(bars + 14);          // up to 64 output channels

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int acc_cols = p.NT;                                  // columns per accumulator
    const uint32_t tmem_cols = 2 * p.NT <= 128 ? 128u : (2 * p.NT <= 256 ? 256u : 512u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tmem_full_bar[s]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[s]), 128);
        }
        mbar_init(smem_u32(w_bar), 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_ptr_smem), tmem_cols);
        tmem_relinquish();
    }
    pdl_wait();          // barrier init / TMEM allocation above overlap the predecessor's tail (launch_k, sscg_common.cuh)
    pdl_launch();
    if (warp >= 2 && warp < 6 && p.bias != nullptr) {
        const int e = threadIdx.x - 64;
        if (e < p.n_ntiles * p.CoW && e < 64) bias_sm[e] = p.bias[e];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int o_first = 3 * p.Wp;       // flattened index of the first output covered by tile 0 of a sample

    if (warp == 0) {
        // ================================ TMA producer ==========================================
        if (lane == 0) {
            int stage = 0; uint32_t par = 0;
            const uint32_t tx = (uint32_t)(p.wstat ? a_bytes : a_bytes + b_bytes);
            if (p.wstat) {      // this CTA's N tile is blockIdx.x % n_ntiles for every tile it walks: load its slabs once
                const int nt0 = blockIdx.x % p.n_ntiles;
                const uint32_t wb = smem_u32(w_bar);
                mbar_arrive_expect_tx(wb, (uint32_t)(7 * b_bytes));
                for (int kh = 0; kh < 7; ++kh)
                    tma_load_2d(smem_u32(wres + kh * b_al), &tmB, wb, 0, (kh * p.n_ntiles + nt0) * p.NT);
            }
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_ntiles;
                const int rest = tile / p.n_ntiles;
                const int t = rest % p.tiles_per_sample, n = rest / p.tiles_per_sample;
                const int px0 = n * p.Hp * p.Wp + o_first + t * kNxOut - 3;      // input pixel of tile row 0, kh = 3
                for (int kh = 0; kh < 7; ++kh) {
                    mbar_wait(smem_u32(&empty_bar[stage]), par ^ 1, 21);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    mbar_arrive_expect_tx(fb, tx);
                    uint8_t* st = smem + stage * stage_bytes;
                    tma_load_2d(smem_u32(st), &tmA, fb, 0, px0 + (kh - 3) * p.Wp);
                    if (!p.wstat) tma_load_2d(smem_u32(st + b_off), &tmB, fb, 0, (kh * p.n_ntiles + nt) * p.NT);
                    if (++stage == p.stages) { stage = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ============================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(kNxTileM, p.NT, 0, 0);
            int stage = 0; uint32_t par = 0;
            uint32_t it = 0;
            if (p.wstat) {
                mbar_wait(smem_u32(w_bar), 0, 25);
                tc_fence_after();
            }
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const uint32_t acc = it & 1, acc_par = (it >> 1) & 1;
                ++it;
                mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_par ^ 1, 22);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * acc_cols;
                uint32_t accum = 0;
                for (int kh = 0; kh < 7; ++kh) {
                    mbar_wait(smem_u32(&full_bar[stage]), par, 23);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                    const uint64_t da = nexp_desc(sa, p.row_bytes);
                    const uint64_t db = nexp_desc(p.wstat ? smem_u32(wres + kh * b_al) : sa + b_off, p.row_bytes);
                    for (int k = 0; k < p.ksteps; ++k) {
                        umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                        accum = 1;
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));
                    if (++stage == p.stages) { stage = 0; par ^= 1; }
                }
                umma_commit(smem_u32(&tmem_full_bar[acc]));
            }
        }
    } else {
        // ================================ epilogue ==============================================
        const int quad = warp & 3;
        const int m = quad * 32 + lane;            // TMEM lane = tile pixel
        const int grp = (warp - 2) >> 2;           // epilogue group 0 (warps 2..5) / 1 (warps 6..9)
        if (grp >= p.groups) goto done;            // a single staging buffer fits: the second group idles
        const uint32_t s_base = smem_u32(S) + (uint32_t)(grp * p.s_bytes);
        const uint32_t bias_s = smem_u32(bias_sm);
        const uint32_t srow = s_base + (uint32_t)(m * s_pitch) * 4u;
        const uint32_t bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;     // named barriers of this group
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_ntiles;
            const int rest = tile / p.n_ntiles;
            const int t = rest % p.tiles_per_sample, n = rest / p.tiles_per_sample;
            const uint32_t acc = it & 1, acc_par = (it >> 1) & 1;
            ++it;
            if (p.groups == 2 && (int)acc != grp) continue;         // with two groups, group g drains accumulator g
            mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_par, 24);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + acc * acc_cols + (static_cast<uint32_t>(quad * 32) << 16);
            const int o = o_first + t * kNxOut + m;               // flattened padded position of this thread's output
            const int orow = o / p.Wp, ocol = o - orow * p.Wp;
            const bool valid = (m < kNxOut) && orow >= 3 && orow < p.Hp - 3 && ocol >= 3 && ocol < p.Wp - 3;
            const long long yoff = (long long)n * p.y_sN + (long long)(orow - 3) * p.y_sH +
                                   (long long)(ocol - 3) * p.y_sW + nt * p.y_c0_step;
            for (int hf = 0; hf < p.halves; ++hf) {
                // ---- accumulator -> staging buffer (row m) --------------------------------------------------
                if (p.halves == 1) {        // all NT columns, two 32-column TMEM loads in flight
                    for (int c0 = 0; c0 < p.NT; c0 += 64) {
                        uint32_t r0[32], r1[32];
                        const int w0 = p.NT - c0 >= 32 ? 32 : 16;
                        const int rem = p.NT - c0 - 32;
                        const int w1 = rem >= 32 ? 32 : (rem >= 16 ? 16 : 0);
                        if (w0 == 32) tmem_ld_32x32(t_acc + c0, r0);
                        else tmem_ld_32x16(t_acc + c0, r0);
                        if (w1 == 32) tmem_ld_32x32(t_acc + c0 + 32, r1);
                        else if (w1 == 16) tmem_ld_32x16(t_acc + c0 + 32, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (q < w0) sts_f32(srow + (uint32_t)(c0 + q) * 4u, __uint_as_float(r0[q]));
#pragma unroll
                        for (int q = 0; q < 32; ++q)
                            if (q < w1) sts_f32(srow + (uint32_t)(c0 + 32 + q) * 4u, __uint_as_float(r1[q]));
                    }
                } else {                    // the ch_half channels of this pass from each of the seven column groups
                    for (int kw = 0; kw < 7; ++kw) {
                        uint32_t r0[32];
                        const uint32_t col = (uint32_t)(kw * p.CoW + hf * ch_half);
                        if (ch_half == 16) tmem_ld_32x16(t_acc + col, r0);
                        else tmem_ld_32x8(t_acc + col, r0);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            if (q < ch_half) sts_f32(srow + (uint32_t)(kw * ch_half + q) * 4u, __uint_as_float(r0[q]));
                    }
                }
                if (hf == p.halves - 1) {
                    tc_fence_before();
                    mbar_arrive(smem_u32(&tmem_empty_bar[acc]));      // accumulator drained
                }
                named_bar_sync(bar_a, 128);                            // every row of S is written
                // ---- shift-add over the seven horizontal taps + bias / activation + store ----------------
                if (valid) {
                    for (int c0 = 0; c0 < ch_half && hf * ch_half + c0 < p.c_store; c0 += 8) {
                        // all 56 staged values first (the loads are ordered volatile asm: adding as they arrive would
                        // expose one shared-memory latency per value), then the sums in tap order
                        float tv[7][8];
#pragma unroll
                        for (int kw = 0; kw < 7; ++kw) {
                            const uint32_t src = s_base + (uint32_t)((m + kw) * s_pitch + kw * ch_half + c0) * 4u;
#pragma unroll
                            for (int q = 0; q < 8; ++q) tv[kw][q] = lds_f32(src + 4u * q);
                        }
                        float v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = 0.f;
#pragma unroll
                        for (int kw = 0; kw < 7; ++kw)
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] += tv[kw][q];
                        const int cch = hf * ch_half + c0;             // first output channel of this vector
                        if (p.bias != nullptr) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] += lds_f32(bias_s + 4u * (uint32_t)(nt * p.y_c0_step + cch + q));
                        }
                        if (p.act == SSCG_ACT_TANH) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] = tanhf(v[q]);
                        } else if (p.act == SSCG_ACT_RELU) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
                        }
                        if (p.y_fp32) {
                            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + yoff + cch);
                            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
                            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yoff + cch);
                            *dst = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                              pack_bf16x2(v[6], v[7]));
                        }
                    }
                }
                named_bar_sync(bar_b, 128);                            // S may be overwritten by the next pass / tile
            }
        }
        tc_fence_before();
    }
done:
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

static int encode_2d_pitch(CUtensorMap* tm, const void* ptr, long long rows, long long pitch_elems, int box_cols,
                           int box_rows, long long dim0 = 0) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)(dim0 > 0 ? dim0 : box_cols), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
    const CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint32_t b[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t s[2] = {1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15))
        return set_error("conv7_nexp: activation pointer/pitch must be 16-byte aligned");
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, b, s,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error("conv7_nexp: cuTensorMapEncodeTiled failed: %d rows=%lld pitch=%lld", (int)r, rows, pitch_elems);
    return 0;
}

}  // namespace sscg

using namespace sscg;

extern "C" int sscg_conv7_nexp(const SscgConv7Args* a, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (a->ksteps != 1 && a->ksteps != 2 && a->ksteps != 4) return set_error("conv7_nexp: ksteps must be 1, 2 or 4");
    if (a->x_pitch < 16 * a->ksteps || a->x_pitch % 8) return set_error("conv7_nexp: bad activation pitch %d", a->x_pitch);
    if (a->CoW % 8 || a->CoW < 8 || a->CoW > 32) return set_error("conv7_nexp: CoW=%d must be 8, 16, 24 or 32", a->CoW);
    if (a->c_store % 8 || a->c_store > a->CoW) return set_error("conv7_nexp: c_store=%d", a->c_store);
    if (a->Hp < 7 || a->Wp < 7 || a->N < 1 || a->n_ntiles < 1) return set_error("conv7_nexp: bad geometry");
    if (a->bias && a->n_ntiles * a->CoW > 64) return set_error("conv7_nexp: bias supports up to 64 output channels");
    NexpDev d;
    d.N = a->N; d.Hp = a->Hp; d.Wp = a->Wp;
    d.CoW = a->CoW;
    d.NT = ((7 * a->CoW + 15) / 16) * 16;
    d.ksteps = a->ksteps;
    d.n_ntiles = a->n_ntiles;
    const long long outs = (long long)(a->Hp - 6) * a->Wp;          // flattened output positions, rows 3 .. Hp-4
    d.tiles_per_sample = (int)((outs + kNxOut - 1) / kNxOut);
    d.total_tiles = d.N * d.tiles_per_sample * d.n_ntiles;
    d.y = a->y; d.y_fp32 = a->y_fp32;
    d.y_sN = a->y_sN; d.y_sH = a->y_sH; d.y_sW = a->y_sW;
    d.y_c0_step = a->CoW;
    d.c_store = a->c_store;
    d.bias = a->bias; d.act = a->act;
    d.row_bytes = 32 * a->ksteps;
    const int a_bytes = kNxTileM * d.row_bytes, b_bytes = d.NT * d.row_bytes;
    const int a_al = (a_bytes + 1023) & ~1023, b_al = (b_bytes + 1023) & ~1023;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = sms < d.total_tiles ? sms : d.total_tiles;
    const int budget = 225 * 1024 - 1024;
    auto s_bytes_of = [&](int halves) {
        const int cols = halves == 1 ? d.NT : 7 * (a->CoW / halves);
        return ((kNxTileM * (cols + 1) * 4 + 15) & ~15) + 16;
    };
    // Shared-memory plan, in order of preference:
    //   two epilogue groups (one staging buffer each) + weight-stationary slabs (the seven slabs of this CTA's N tile stay
    //   resident: they were re-streamed for every 122-pixel tile, 100 of 156 KB for the head's data gradient) + >= 3 stages,
    //   then two groups without resident weights, then one group.  The staging buffer shrinks by running the shift-add
    //   in `halves` passes of CoW / halves channels (a multiple of 8).
    const int pass_opts[4] = {1, 2, 3, 4};
    d.wstat = 0; d.halves = 1; d.groups = 1;
    bool planned = false;
    const int grid_w = grid - grid % d.n_ntiles;                 // every CTA keeps one N tile
    const bool allow_wstat = !getenv("SSCG_NEXP_NO_WSTAT") && grid_w >= d.n_ntiles;
    const int max_groups = getenv("SSCG_NEXP_ONE_GROUP") ? 1 : 2;
    // (measured at 16 x 256 x 256: more than two passes cost more than a second epilogue group or resident weights
    //  gain — 437 us with 2 groups / 4 passes against 366 us with 1 group / 2 passes for K = 32 x 7, N = 224 x 2; two
    //  groups + resident weights + 2 passes: 369 -> 318 us for K = 16 x 7)
    for (int max_pass = 2; max_pass <= 4 && !planned; max_pass += 2)
        for (int ws = allow_wstat ? 1 : 0; ws >= 0 && !planned; --ws)
            for (int groups = max_groups; groups >= 1 && !planned; --groups)
                for (int pi = 0; pi < 4 && !planned; ++pi) {
                    const int h = pass_opts[pi];
                    if (h > max_pass || a->CoW % h || (a->CoW / h) % 8 || (h > 1 && (a->CoW / h) > 16)) continue;
                    const int need = groups * s_bytes_of(h) + (ws ? 7 * b_al + 3 * a_al : 2 * (a_al + b_al));
                    if (need <= budget) {
                        d.groups = groups; d.wstat = ws; d.halves = h;
                        planned = true;
                    }
                }
    if (!planned) return set_error("conv7_nexp: tile does not fit shared memory (NT=%d)", d.NT);
    if (d.wstat) grid = grid_w;
    d.s_bytes = s_bytes_of(d.halves);
    const int stage_bytes = d.wstat ? a_al : a_al + b_al;
    const int fixed = d.groups * d.s_bytes + (d.wstat ? 7 * b_al : 0) + 512;
    int stages = (budget - fixed) / stage_bytes;
    if (stages > 4) stages = 4;
    if (stages < 2) return set_error("conv7_nexp: tile does not fit shared memory (NT=%d)", d.NT);
    d.stages = stages;
    const int smem = 1024 + stages * stage_bytes + fixed;
    if (getenv("SSCG_DEBUG"))
        fprintf(stderr, "[sscg] conv7_nexp NT=%d K=%dx7: groups %d, wstat %d, passes %d, stages %d, smem %d\n", d.NT,
                16 * a->ksteps, d.groups, d.wstat, d.halves, stages, smem);

    CUtensorMap tmA, tmB;
    const long long rows = (long long)a->N * a->Hp * a->Wp;
    const int kcols = 16 * a->ksteps;
    if (int rc = encode_2d_pitch(&tmA, a->x, rows, a->x_pitch, kcols, kNxTileM)) return rc;
    if (int rc = encode_2d_pitch(&tmB, a->w, 7LL * d.n_ntiles * d.NT, 64, kcols, d.NT, 64)) return rc;

    static int max_smem_set = 0;
    if (smem > max_smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_nexp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return set_error("conv7_nexp: cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
        max_smem_set = smem;
    }
    {
        LaunchScope ls(a->tag, stream);
        launch_k(conv_nexp_kernel, grid, 320, smem, stream, tmA, tmB, d);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("conv7_nexp launch: %s", cudaGetErrorString(e));
    return 0;
}
