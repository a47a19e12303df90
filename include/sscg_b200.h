/* sscg_b200.h — C ABI of libsscg_b200.so: the B200 (sm_100a) kernels behind the reference's
 * network boundary `arch.define_Gen / define_Dis / set_grad` (reference arch/__init__.py:1-3).
 *
 * The reference has no FFI of its own (it is pure Python over torch.nn); each entry point below
 * replaces one family of ATen/cuDNN calls that the reference's modules issue on the hot path
 * (SURVEY.md §2.3 K1-K19).  Plain pointers and sizes only: every pointer is a DEVICE pointer
 * unless its name ends in `_host`; every function enqueues on `stream` (a cudaStream_t passed as
 * void*) and returns immediately.  Return value: 0 on success, non-zero on failure — the message
 * is available from sscg_last_error().  No hidden allocation, no global mutable state besides the
 * last-error string and the lazily resolved driver entry point for tensor-map encoding.
 *
 * Layout vocabulary: activations are NHWC bf16 ("planes" of one sample are H x W x Cpitch); a
 * "view" is (ptr, N, H, W, C, sN, sH, sW) with element strides, which lets one buffer carry an
 * explicit reflect halo while another kernel reads only its interior (zero fill outside).
 */
#ifndef SSCG_B200_H
#define SSCG_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSCG_MAX_TAPS 64

/* activation codes */
#define SSCG_ACT_NONE 0
#define SSCG_ACT_RELU 1
#define SSCG_ACT_LRELU 2
#define SSCG_ACT_TANH 3

/* Plane sums (InstanceNorm statistics and their backward counterparts) are kept as BINNED FIXED-POINT accumulators:
 * a value v is the pair of int64 words (hi, lo) with v = hi * 2^-8 + lo * 2^-56.  Kernels add float partial sums with
 * 64-bit integer atomics (exact split of the float into the two words), so the total does not depend on the order in
 * which thread blocks arrive: results are reproducible bit for bit, with no fence, counter or extra launch.  A buffer
 * of plane sums is int64 [N][C][2][SSCG_STAT_WORDS] and must be zeroed before the launch that accumulates into it. */
#define SSCG_STAT_WORDS 2

/* halo modes */
#define SSCG_PAD_NONE 0
#define SSCG_PAD_ZERO 1
#define SSCG_PAD_REFLECT 2

typedef struct SscgTap {
    int8_t dh, dw;  /* input pixel = out pixel * stride + (dh, dw) + (org_h, org_w) */
    int16_t brow;   /* weight slab index: weight rows [brow*Co_pad, (brow+1)*Co_pad) */
} SscgTap;

/* NHWC bf16 view (element strides). C is the extent of the innermost (contiguous) dimension that a
 * K-block may address: the channel pitch in regular mode, the padded (kw,c) window in window mode. */
typedef struct SscgView {
    const void* ptr;
    int32_t N, H, W, C;
    int64_t sN, sH, sW;
} SscgView;

/* ---------------------------------------------------------------------------------------------
 * sscg_conv_igemm — implicit-GEMM convolution on tcgen05 tensor cores (TMA-fed, TMEM accumulate).
 * Replaces: nn.Conv2d / nn.ConvTranspose2d forward (reference arch/ops.py:40-57,63,68;
 * arch/generators.py:74-90; arch/discriminators.py:45-58) and their cuDNN dgrad (K17), with the
 * reflection pad (arch/ops.py:62,67; generators.py:73,84,89) folded into the operand view and the
 * InstanceNorm statistics (arch/ops.py:11) accumulated in the epilogue.
 *
 *   Y[n, ho, wo, co] = act( bias[co] + sum_{tap, k} X[n, ho*stride+tap.dh+org_h, wo*stride+tap.dw+org_w, k]
 *                                                   * Wt[tap.brow*Co_pad + co, k] )
 * Pixels outside the X view read as zero.  With n_phases == 4 the output is produced as four
 * interleaved sub-grids (ho = 2i+ph, wo = 2j+pw; phase p = 2*ph+pw uses taps
 * [phase_start[p], phase_start[p+1])) — the sub-pixel form of a stride-2 transposed convolution.
 * split == 3 selects the bf16x3 parity mode: X and Wt are given as hi/lo bf16 planes and the
 * product is accumulated as hi*hi + lo*hi + hi*lo in fp32.
 * ------------------------------------------------------------------------------------------- */
typedef struct SscgConvArgs {
    SscgView x;            /* activation view (hi plane) */
    const void* x_lo;      /* lo plane (same strides) or NULL */
    int32_t stride;        /* 1 or 2 */
    int32_t Kc;            /* K elements per tap; multiple of 64; <= x.C */
    int32_t org_h, org_w;
    int32_t n_phases;      /* 1 or 4 */
    int32_t phase_start[5];
    SscgTap taps[SSCG_MAX_TAPS];
    const void* w;         /* bf16 [w_rows][Kc] (K-major), hi plane */
    const void* w_lo;      /* lo plane or NULL */
    int32_t w_rows;
    int32_t Co_pad;        /* output channels incl. padding; multiple of BN */
    int32_t split;         /* 1 (bf16) or 3 (bf16x3) */
    void* y;               /* output: bf16 (y_fp32 == 0) or fp32 (y_fp32 == 1), NHWC-like */
    int32_t y_fp32;
    int64_t y_sN, y_sH, y_sW;   /* element strides; channel stride is 1 */
    int32_t y_oh, y_ow;    /* output written at (ho + y_oh, wo + y_ow) */
    int32_t Ho, Wo;        /* output extents (all phases together) */
    const float* bias;     /* [Co_pad] or NULL */
    int32_t act;           /* SSCG_ACT_* applied in the epilogue */
    float slope;
    void* stats;           /* int64 [N][Co_pad][2][SSCG_STAT_WORDS]: binned accumulators of (sum, sum of squares) per
                            * (sample, channel) plane, ADDED to by the launch (zero them first); or NULL.  See
                            * "Plane sums" below. */
    int32_t TH, TW;        /* output tile, TH*TW == 128 */
    int32_t BN;            /* N tile: 16, 32, 64, 128 or 256 */
    int32_t tag;           /* profiling class (0..15), see sscg_prof_begin */
    /* Row-shift mode (0 = off): for stride-1 convolutions with shift_kw (= 7) horizontal taps per filter
     * row, `taps` holds ONE entry per filter row (dh, leftmost dw, slab index of the leftmost tap); the
     * kernel loads a (TW + shift_kw - 1)-pixel row box once and feeds the shift_kw taps from shifted
     * shared-memory descriptors; tap j uses weight slab brow + j * shift_brow_step.  Needs TH = 1,
     * TW = 128, BN in {16, 32, 64}, split == 1. */
    int32_t shift_kw;
    int32_t shift_brow_step;
    int32_t shift_base_mode;   /* 2 (use this): descriptor base_offset 0 — the 128B swizzle is a pure function of the
                                * shared-memory address (verified on B200, tools/shift_probe.py); 1: base_offset = row
                                * phase (diagnostic only: produces wrong results) */
    /* Flattened tiling (stride-1 gathers over a buffer that carries its zero halo explicitly): x is the 1-D pixel
     * view (N = H = 1, W = samples * flat_hw) of a [flat_n][rows][flat_pitch][C] buffer, a tile is 128 consecutive
     * positions f = row * flat_pitch + col of one sample, tap (dh, dw) reads position f + dh * flat_pitch + dw, and
     * positions with col >= Wo or row >= Ho are computed but not stored.  66 x 66 outputs then take 36 tiles per
     * sample instead of 45 (8 x 16 tiles).  0 = off.  Requires stride 1, one phase, TH = 1, TW = 128. */
    int32_t flat_pitch, flat_hw, flat_n;
    /* Pixel-row mode for stems (0 = off, else bytes per pixel of x: 16 * G for x.C == 8 * G channels, G = 1..4): `taps`
     * holds ONE entry per filter row, Kc == 64 * G, weights in sscg_wprep mode 5 ([kh][Co_pad][(g, kw, c8)]), x is the PLAIN
     * haloed view; per filter row and channel group the kernel loads one dense (128 + 8)-pixel x 8-channel row box and
     * addresses it as overlapping K-major rows.  Needs stride 1, one phase, TH = 1, TW = 128, BN = 64, split == 1. */
    int32_t rw_pitch;
} SscgConvArgs;

int sscg_conv_igemm(const SscgConvArgs* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * sscg_conv_wgrad — weight gradient as a pixel-contraction GEMM on tcgen05 (both operands MN-major).
 * Replaces: cuDNN wgrad of the same convolutions (K17).
 *   dWt[tap.brow*Co_pad + co, k] += sum_{n, ho, wo} dY[n, ho, wo, co] * X[n, ho*stride+tap.dh+org_h, ..., k]
 * dWt is fp32 [w_rows][Kc].  Split-K over samples / pixel tiles: the ksplit partial tiles of an output tile go to
 * the workspace and are summed in split order by the CTA that arrives last, which alone adds the sum to dWt
 * (no floating-point atomics: results do not depend on CTA scheduling).
 * ------------------------------------------------------------------------------------------- */
typedef struct SscgWgradArgs {
    SscgView dy;           /* output-gradient view [N][Ho][Wo][Co_pad] (hi plane); zero outside */
    const void* dy_lo;
    SscgView x;            /* forward activation view (hi plane) */
    const void* x_lo;
    int32_t stride, Kc, org_h, org_w;
    int32_t n_taps;
    SscgTap taps[SSCG_MAX_TAPS];
    int32_t Co_pad;        /* multiple of 64 */
    int32_t split;
    float* dw;             /* fp32 [w_rows][Kc] */
    int32_t w_rows;
    int32_t TH, TW;        /* pixel block, TH*TW == 64 */
    int32_t BN;            /* K-column tile of dWt: 64, 128 or 256 (divides Kc) */
    int32_t ksplit;        /* number of CTAs sharing one (tap, co-tile, k-tile) */
    int32_t tag;           /* profiling class (0..15) */
    void* ws;              /* ksplit > 1: workspace of sscg_conv_wgrad_ws_bytes() bytes (arrival counters, zero before
                            * the first launch and left zero by every launch, + partial tiles) */
    int32_t rw_pitch;      /* pixel-row mode for stride-1 stems (0 = off, else 16 * G for x.C == 8 * G channels): x is the
                            * PLAIN haloed view, `taps` holds one entry per filter row, dWt columns are in sscg_wprep mode 5
                            * order (group, kw, channel in group); needs TH = 1, TW = 64, BN = Kc = 64 * G, split == 1 */
} SscgWgradArgs;

int sscg_conv_wgrad(const SscgWgradArgs* a, void* stream);
int64_t sscg_conv_wgrad_ws_bytes(const SscgWgradArgs* a);
/* CTAs of this tile shape that share an SM (1 or 2): size ksplit so that the grid fills ctas_per_sm * #SM slots. */
int32_t sscg_conv_wgrad_ctas_per_sm(int32_t BN, int32_t split);

/* sscg_conv_wgrad7 — weight gradient of the 7x7 stride-1 generator HEAD (64 -> <= 32 channels, arch/generators.py:84-85,
 * 89-90) with the seven horizontal taps as GEMM columns (csrc/conv_wgrad7.cu): replaces the window-mode launch of
 * sscg_conv_wgrad for that layer (631 -> ~100 us at 16 x 256 x 256).
 *   x : bf16 [N][H+6][W+6][64]   the head's input activation with its explicit (reflect) halo of 3
 *   dy: bf16 [N][H+12][W+12][Cy] the gradient w.r.t. the head's raw output in a buffer with a ZERO halo of 6 (the layout
 *       sscg_conv7_nexp's data gradient reads); Cy = 16 or 32 channels per pixel, padding channels zero
 *   dw: fp32 [7 kh][64 rows = co][448 = (kw, ci)], ADDED to — the window-mode slab layout (sscg_wprep mode 1)
 *   ws: sscg_conv_wgrad7_ws_bytes() bytes (counters zero before the first launch, left zero by every launch); the
 *       per-CTA tiles are combined in CTA order: reproducible. */
typedef struct SscgWgrad7Args {
    const void* x;
    const void* dy;
    int32_t N, H, W;
    int32_t Cy;
    float* dw;
    void* ws;
    int32_t tag;
} SscgWgrad7Args;
int sscg_conv_wgrad7(const SscgWgrad7Args* a, void* stream);
int64_t sscg_conv_wgrad7_ws_bytes(const SscgWgrad7Args* a);

/* ---------------------------------------------------------------------------------------------
 * Elementwise / reduction kernels around the GEMMs.
 * ------------------------------------------------------------------------------------------- */

/* sscg_pack_nchw: NCHW fp32 -> NHWC bf16 with channel padding and an explicit halo.
 * Replaces the layout change + nn.ReflectionPad2d(3) at the generator stem (generators.py:73) and
 * the zero padding of the PatchGAN stem (discriminators.py:45).  dst extents (H+2*pad, W+2*pad). */
int sscg_pack_nchw(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, void* dst, void* dst_lo,
                   int32_t dst_fp32, int32_t Cp, int32_t pad, int32_t pad_mode, void* stream);

/* sscg_onehot_pack: int64 label map [N][1][H][W] -> one-hot NHWC bf16 with halo (utils.py:314-350
 * make_one_hot fused with the stem's layout change). */
int sscg_onehot_pack(const int64_t* labels, int32_t N, int32_t C, int32_t H, int32_t W, void* dst, void* dst_lo,
                     int32_t Cp, int32_t pad, int32_t pad_mode, void* stream);

/* sscg_unpack_nhwc: NHWC fp32 [N][H][W][Cp] -> NCHW fp32 [N][C][H][W] (module output boundary). */
int sscg_unpack_nhwc(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cp, float* dst,
                     void* stream);

/* sscg_unpack_fold: gradient of the boundary pack — padded NHWC (bf16 or fp32) [N][H+2p][W+2p][Cp] ->
 * NCHW fp32 [N][C][H][W], folding halo gradients back onto their source pixels (backward of
 * nn.ReflectionPad2d / zero padding at generators.py:73, discriminators.py:45). */
int sscg_unpack_fold(const void* src, int32_t src_fp32, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cp,
                     int32_t pad, int32_t pad_mode, float* dst, void* stream);

/* sscg_bias_grad: grad[c] += scale * sum_n bstats[n][c][0] (bias gradient of a conv that is not
 * followed by InstanceNorm: head conv generators.py:85,90; PatchGAN stem/tail discriminators.py:45,58);
 * bstats in the plane-sum format. */
int sscg_bias_grad(const void* bstats, int32_t N, int32_t C, int32_t Cp, float* grad, float scale, void* stream);

/* sscg_in_apply: y = dropout(act(instance_norm(raw))) (+ residual), written with a halo for the next
 * convolution.  Replaces nn.InstanceNorm2d + ReLU/LeakyReLU + Dropout + residual add + the next
 * layer's ReflectionPad2d (ops.py:11,44,50,57,62-74).
 *   raw:   [N][H][W][C] bf16 (raw_fp32 == 0) or fp32
 *   stats: plane sums [N][C][2][SSCG_STAT_WORDS] (sum, sumsq) from the conv epilogue, or NULL for "no norm"
 *   res:   optional residual view (bf16 hi/lo), added after norm/act
 *   dst:   bf16 [N][H+2p][W+2p][C] (+ lo plane when dst_lo != NULL)
 *   dropout: p = 0.5 when drop_seed != 0 (keep mask = hash(seed, element index); scale 2). */
typedef struct SscgApplyArgs {
    const void* raw; int32_t raw_fp32;
    const void* stats; float eps;
    int32_t N, H, W, C;
    int32_t act; float slope;
    uint64_t drop_seed;
    const uint64_t* drop_ctr;   /* optional device counter mixed into the seed (CUDA-graph replays) */
    SscgView res; const void* res_lo;
    void* dst; void* dst_lo;
    int32_t pad, pad_mode;
} SscgApplyArgs;
int sscg_in_apply(const SscgApplyArgs* a, void* stream);

/* sscg_in_bwd_prep / sscg_in_bwd_apply: backward of the same chain.
 * prep:  dZ = act'(Z) * dropmask * ( fold_halo(dYp) + skip ), Z = instance_norm(raw) recomputed;
 *        writes dZ (and optionally the folded sum G for the residual skip path), accumulates
 *        bstats[n][c] = (sum dZ, sum dZ*Z) (plane-sum format).  With stats == NULL (no norm) dZ is the final dRaw and
 *        bstats[.][c][0] is the bias gradient contribution.
 * apply: dRaw = rstd * (dZ - mean(dZ) - Z * mean(dZ*Z)). */
typedef struct SscgBwdArgs {
    const void* raw; int32_t raw_fp32;
    const void* stats; float eps;
    int32_t N, H, W, C;
    int32_t act; float slope;
    uint64_t drop_seed;
    const uint64_t* drop_ctr;
    SscgView dyp; int32_t dyp_fp32;   /* gradient w.r.t. the padded consumer buffer; interior offset = pad */
    int32_t pad, pad_mode;
    SscgView skip; int32_t skip_fp32; /* optional extra gradient on the unpadded output (residual path) */
    void* g_out; int32_t g_fp32;      /* optional: folded dYp + skip, [N][H][W][C] */
    void* dz; int32_t dz_fp32;        /* [N][H][W][C] */
    void* dz_lo;                      /* lo plane when dZ is the final dRaw in split mode */
    void* bstats;                     /* plane sums [N][C][2][SSCG_STAT_WORDS] of (dZ, dZ*Z), ADDED to (zero them first) */
    int32_t dz_pad;                   /* > 0: dZ is written into a [N][H+2p][W+2p][C] buffer at offset (p, p); the halo is
                                         left untouched (kept zero by the caller: input layout of sscg_conv7_nexp's data gradient) */
    int32_t draw_pad;                 /* same, for the dRaw output of sscg_in_bwd_apply */
} SscgBwdArgs;
int sscg_in_bwd_prep(const SscgBwdArgs* a, void* stream);
int sscg_in_bwd_apply(const SscgBwdArgs* a, void* draw, void* draw_lo, void* stream);
/* sscg_set_stream_norm: 2 (default) lets sscg_in_apply / sscg_in_bwd_prep / sscg_in_bwd_apply use their
 * bulk-copy pipelined variants (cp.async.bulk ring in shared memory, one persistent CTA per SM, a producer
 * warp and free-running consumer warps) when every tensor is bf16 and the row geometry suits the ring; 1
 * keeps the register-batched sscg_in_bwd_prep; 0 forces the register-batched kernels everywhere.  Same arithmetic in every mode; the
 * summation order of the plane sums is fixed within a mode and differs between modes.  SSCG_STREAM_NORM=0|1|2 sets the initial value. */
int sscg_set_stream_norm(int32_t on);

/* weight preparation: fp32 master weights -> bf16 GEMM operand slabs (see DESIGN.md "weight slabs") */
typedef struct SscgWprepArgs {
    const float* w;        /* Conv2d: [Co][Ci][KH][KW]; ConvTranspose2d: [Ci][Co][KH][KW] */
    int32_t transposed;    /* source is ConvTranspose2d layout */
    int32_t Co, Ci, KH, KW;
    int32_t mode;          /* 0: fwd regular  [tap][Co_pad][Kc(ci)]
                              1: fwd window   [kh][Co_pad][Kc(kw*Cp+ci)]
                              2: dgrad regular [tap][Ci_pad][Kc(co)]  (taps indexed as in the source)
                              3: N-expanded 7x7 forward  [kh][nt][NT: kw * CoW + (co - nt * CoW)][Kc(ci)]
                              4: N-expanded 7x7 dgrad    [kh][nt][NT: kw * CoW + (ci - nt * CoW)][Kc(co)], taps flipped
                                 (sscg_conv7_nexp; CoW is passed in Cp, rows_pad = NT = round_up(7 * CoW, 16))
                              5: fwd pixel-row [kh][Co_pad][Kc: 64 * g + 8 * kw + (ci - 8 * g)], channel groups g of 8
                                 (sscg_conv_igemm rw_pitch; Kc = 64 * Cp / 8, KW <= 8)
                            */
    int32_t Cp;            /* channel pitch of the activation in window mode; CoW in modes 3 / 4 */
    int32_t rows_pad;      /* Co_pad (mode 0,1) or Ci_pad (mode 2) */
    int32_t Kc;
    void* dst; void* dst_lo;
} SscgWprepArgs;
int sscg_wprep(const SscgWprepArgs* a, void* stream);
/* inverse mapping for gradients: fp32 slab [rows][Kc] -> += into the parameter-shaped gradient */
int sscg_wgrad_unpack(const SscgWprepArgs* a, const float* slab, float* grad, float scale, void* stream);

/* sscg_conv7_nexp: 7x7 stride-1 convolution with a narrow output (generator head 64 -> 21 / 3 channels,
 * arch/generators.py:84-85,89-90, and its data gradient) as an N-expanded implicit GEMM: the seven horizontal
 * taps become GEMM columns (N = 7 * CoW) and are shift-added in the epilogue through shared memory, so a
 * 128-pixel tile needs 7 * ksteps wide MMAs instead of 196 narrow ones (see csrc/conv_nexp.cu).
 *   x: bf16 activation [N][Hp][Wp][x_pitch] carrying an explicit halo of 3 (forward: the reflect halo of the
 *      head's input; data gradient: dRaw in a buffer with a ZERO halo of 6); the first 16 * ksteps channels of
 *      every pixel enter the contraction.
 *   w: slab from sscg_wprep mode 3 (forward) or 4 (data gradient): [7][n_ntiles][NT][64] bf16.
 *   y: output [N][Hp-6][Wp-6] with the given element strides, fp32 (y_fp32) or bf16; N tile nt writes channels
 *      nt * CoW .. nt * CoW + c_store; bias (fp32, indexed by output channel) and activation are optional. */
typedef struct SscgConv7Args {
    const void* x; int32_t x_pitch;
    int32_t N, Hp, Wp;
    const void* w;
    int32_t CoW, n_ntiles, ksteps, c_store;
    void* y; int32_t y_fp32;
    int64_t y_sN, y_sH, y_sW;
    const float* bias; int32_t act;
    int32_t tag;
} SscgConv7Args;
int sscg_conv7_nexp(const SscgConv7Args* a, void* stream);

/* Batched forms: ONE launch over a DEVICE table of descriptors (every slab of a network).  `start` is the
 * running (row, k) offset of the entry (entries sorted by start; total = sum of rows_pad * Kc — a thread owns one
 * (row, k) position and walks the taps);
 * slab / grad are used by the gradient form only.  Same arithmetic as the per-slab calls above. */
typedef struct SscgWbatchEntry {
    SscgWprepArgs a;
    const float* slab;
    float* grad;
    int64_t start;
} SscgWbatchEntry;
int sscg_wprep_batch(const SscgWbatchEntry* table_dev, int32_t count, int64_t total, void* stream);
int sscg_wgrad_unpack_batch(const SscgWbatchEntry* table_dev, int32_t count, int64_t total, float scale, void* stream);

/* Loss reductions are grid-wide sums WITHOUT floating-point atomics: every block stores its partial sums to a slot
 * of a caller-provided workspace of SSCG_LOSS_WS_BYTES bytes (zero before the first use; every launch leaves its
 * arrival counter zero) and the block that arrives last adds the slots in a fixed order, so results are
 * reproducible bit for bit.  One workspace must not be shared by launches that may run concurrently. */
#define SSCG_LOSS_WS_BYTES 16384

/* sscg_seg_head_fwd / _bwd: fused segmentation-head loss on NCHW fp32 logits — softmax over classes
 * (nn.Softmax2d, model.py:273,401-402), cross-entropy against the label map (nn.CrossEntropyLoss, model.py:272,
 * 398,455; log-sum-exp form) and first-max argmax (model.py:435,509) in one pass; backward = cross-entropy
 * gradient + softmax Jacobian of an incoming probability gradient.
 *   labels [N][H][W] int64 or NULL.  Pixels labelled ignore_index (torch default -100) are left out of the loss and
 *   of its gradient; any other label outside [0, C) raises the device error flag (sscg_device_error, code 31) —
 *   torch device-asserts on those — and is left out as well.
 *   probs / argmax may be NULL.  loss_out[2] (or NULL) = (sum of -log p[label] over the counted pixels, their
 *   number): the mean cross-entropy is loss_out[0] / loss_out[1].
 *   backward: dloss = device scalar, gradient of the MEAN cross-entropy (or NULL), count = device scalar holding
 *   loss_out[1]; dprobs or NULL. */
int sscg_seg_head_fwd(const float* logits, const int64_t* labels, int32_t N, int32_t C, int64_t HW,
                      int64_t ignore_index, float* probs, int64_t* argmax, float* loss_out, void* ws, void* stream);
int sscg_seg_head_bwd(const float* probs, const int64_t* labels, const float* dloss, const float* count,
                      const float* dprobs, int32_t N, int32_t C, int64_t HW, float* dlogits, void* stream);

/* sscg_lsgan_fwd / _bwd: LSGAN patch loss against a constant target (nn.MSELoss vs all-ones / all-zeros,
 * model.py:270,445-446,452,521-534): *loss_out = scale * sum (x - target)^2 (scale = 1/n gives the mean);
 * dx = dloss * 2 (x - target) / n with dloss a device scalar (gradient of the MEAN). */
int sscg_lsgan_fwd(const float* x, int64_t n, float target, float scale, float* loss_out, void* ws, void* stream);
int sscg_lsgan_bwd(const float* x, int64_t n, float target, const float* dloss, float* dx, void* stream);
/* sscg_l1_fwd / _bwd: nn.L1Loss (model.py:271,453,461): *loss_out = scale * sum |x - y|; dx = dloss * sign(x - y) / n.
 * Any alignment of x / y is accepted (128-bit loads when both are 16-byte aligned). */
int sscg_l1_fwd(const float* x, const float* y, int64_t n, float scale, float* loss_out, void* ws, void* stream);
int sscg_l1_bwd(const float* x, const float* y, int64_t n, const float* dloss, float* dx, void* stream);
/* sscg_adam_flat: one Adam update (torch.optim.Adam semantics, no weight decay / amsgrad; model.py:286-287,
 * 474,542) over a flat fp32 bucket: p, g, m (exp_avg), v (exp_avg_sq) of n elements; lr and step are DEVICE
 * scalars (step already incremented for this update), so the call is CUDA-graph capturable and the
 * LambdaLR schedule (utils.py:434-441) only rewrites one float. */
int sscg_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float beta1, float beta2,
                   float eps, const float* step, void* stream);

/* sscg_interp_bilinear_fwd / _bwd: bilinear resize with align_corners = True on NCHW fp32 — the reference's `interp`
 * = nn.Upsample(size=(crop_height, crop_width), mode='bilinear', align_corners=True) (model.py:62-63,268), applied to
 * every generator output when the generator (deeplab) works at 1/8 resolution (model.py:390-392,413-415).  x [N][C][Hi][Wi]
 * -> y [N][C][Ho][Wo]; backward: dy -> dx as a gather over the output gradient (no atomics, reproducible). */
int sscg_interp_bilinear_fwd(const float* x, int32_t N, int32_t C, int32_t Hi, int32_t Wi, float* y, int32_t Ho, int32_t Wo,
                             void* stream);
int sscg_interp_bilinear_bwd(const float* dy, int32_t N, int32_t C, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo,
                             float* dx, void* stream);

/* sscg_confusion: hist[t * n_class + p] += #{i : label_true[i] == t, label_pred[i] == p, 0 <= t, p < n_class} —
 * the device form of runningScore._fast_hist (utils.py:363-369) used by the validation loop (model.py:555-572). */
int sscg_confusion(const int64_t* label_true, const int64_t* label_pred, int64_t n, int32_t n_class, uint64_t* hist,
                   void* stream);

/* utility */
int sscg_fill_zero(void* ptr, int64_t bytes, void* stream);
const char* sscg_last_error(void);
/* Programmatic dependent launch of the library's kernels (default off; SSCG_PDL=1 in the environment turns it on):
 * a kernel may be scheduled while its predecessor in the stream drains and blocks (griddepcontrol.wait) before it
 * touches global memory. */
int sscg_set_pdl(int32_t on);
int sscg_device_error(void);   /* reads (and clears) the device-side protocol error flag; 0 = none */
int sscg_version(void);

/* Measurement hooks (bench.py): every launcher counts its launch; while profiling is on, launches
 * are bracketed by CUDA events on their own stream and accumulated per `tag`
 * (conv/wgrad: the tag in the argument block; elementwise kernels: 7 = apply, 8 = backward
 * prep/apply, 9 = pack/unpack/weight prep).  sscg_prof_end synchronises the device and returns the
 * summed milliseconds and launch counts per tag (arrays of 16). */
uint64_t sscg_launch_count(void);
int sscg_prof_begin(uint32_t tag_mask);   /* bit t set: record launches of tag t; 0 = all */
int sscg_prof_end(float* sum_ms, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* SSCG_B200_H */
