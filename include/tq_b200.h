/*
 * tq_b200.h -- C ABI of libtq_b200.so: the B200 (sm_100a) fake-quantization hot path.
 *
 * The reference (Qualcomm-AI-research/transformer-quantization) has no native/FFI boundary: its
 * hot path is a chain of ATen ops inside Python classes.  This header is the boundary a native
 * replacement binds to; every entry point names the reference code it replaces (file:line relative
 * to the reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless a parameter says "host".  Tensors are contiguous,
 *     row-major fp32 unless stated.
 *   - `stream` is a cudaStream_t passed as void*.  Every call only enqueues work on that stream:
 *     no allocation, no host synchronisation, no hidden global state -> CUDA-graph capturable.
 *   - Workspaces are caller-owned, must be zero-initialised ONCE when allocated (the kernels leave
 *     them zeroed again), and must not be shared by two streams at the same time.
 *   - Return value: 0 = enqueued; <0 = TQ_E* argument error (nothing enqueued);
 *     >0 = cudaError_t from the launch.
 *   - Quantizer state stays on the device (the reference's `_delta`, `_zero_float`, `_signed`
 *     buffers): kernels derive scale / zero_point / integer grid from it themselves, which removes
 *     the `.item()` host syncs of quantizers.py:311-328.
 */
#ifndef TQ_B200_H
#define TQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ_OK 0
#define TQ_EINVAL (-1)   /* bad argument (null pointer, negative size, n_bits out of range) */
#define TQ_EALIGN (-2)   /* pointer alignment the entry point requires is not met */
#define TQ_EWORKSPACE (-3) /* workspace too small */
#define TQ_EUNSUPPORTED (-4)

/* Quantizer state as the reference keeps it (quantizers.py:101-102, 308): raw `_delta`,
 * `_zero_float` (asymmetric) or `_signed` (symmetric).  The kernels apply
 *   scale      = max(delta, eps)            | exp(delta) for log_domain   (quantizers.py:142-147)
 *   zero_point = clamp(round(zero_float), 0, 2^n-1)   | 0 for symmetric   (quantizers.py:149-153, 330-332)
 *   [int_min, int_max] = [0, 2^n-1]                    asymmetric          (quantizers.py:131-140)
 *                      = [-2^(n-1), 2^(n-1)-1] if *is_signed else [0, 2^n-1]   (quantizers.py:321-328)
 * n_params = 1 (per-tensor) or C (per-axis / per-embedding-group / per-channel).  */
typedef struct tq_qspec {
    const float*   delta;       /* [n_params] */
    const float*   zero_float;  /* [n_params]; NULL => symmetric quantizer */
    const uint8_t* is_signed;   /* [1] (torch.bool); used only when zero_float == NULL */
    int32_t        n_bits;      /* 1..16 */
    int32_t        log_domain;  /* 0: 'linear', 1: 'log' scale_domain */
    float          eps;         /* quantizer.eps, default 1e-8 */
} tq_qspec;

/* ---- library info ------------------------------------------------------------------------- */
int         tq_version(void);              /* ABI version: 1 = inference path; 2 adds the training-time entry points
                                              * (tq_qdq_bwd_f32, tq_adaround_*) and tq_probe_copy_f32; 3 adds
                                              * tq_linear_seg_qdq_i8, tq_linear_nonorm_qdq_i8, tq_calib_finalize_f32,
                                              * tq_attention_pad_qdq_i8 and the tq_*_peg_* entry points; 4 (current) adds the
                                              * encoder chain (tq_chain_plan_create / _run / _destroy), tq_head_qdq_i8 and the packed-path
                                              * counters of tq_selftest_div (uint64[5]) */
const char* tq_error_string(int code);     /* static string for TQ_E* / cudaError_t */
int         tq_device_sm_count(void);      /* SM count of the current device (148 on B200) */

/* Device self-test of the division-free quotient used by every quantizing kernel (csrc/
 * tq_common.cuh, tq::div_rn): blocks*256*iters adversarial (x, scale) pairs are compared bit for
 * bit with the IEEE division instruction.  mismatches[0] += #quotient mismatches with
 * 2^-60 <= |x/s| < 2^22, mismatches[1] += #integer-grid mismatches (both must stay 0),
 * mismatches[2] += #quotient mismatches below 2^-60 (residual underflow; they round to 0 either
 * way); mismatches[3] / [4] += the same two counts for the PACKED forms the fused epilogues use (quot2,
 * quant_int2_finite, quant_ctr2_finite on FFMA2; both must stay 0).  Device uint64[5], caller-zeroed. */
int tq_selftest_div(uint64_t seed, int32_t blocks, int32_t iters, uint64_t* mismatches, void* stream);

/* Copy-bandwidth probe with this library's streaming access pattern (128-bit grid-stride loop of the
 * quant-dequant kernels, no arithmetic): the attainable ceiling of an 8 B / element kernel on the device at
 * hand.  flags: bit 0 = loads bypass L1 (ld.global.L1::no_allocate), bit 1 = stores bypass L1, bit 2 =
 * persistent grid (<= 4 CTAs per SM, grid-stride) instead of one 16 KB chunk per CTA.  n % 4 == 0,
 * 16-byte aligned pointers. */
int tq_probe_copy_f32(const float* x, float* y, int64_t n, int32_t flags, void* stream);

/* ---- quantize -> round -> clamp -> dequantize ------------------------------------------------
 * a1+a2: AsymmetricUniformQuantizer.forward / SymmetricUniformQuantizer (quantizers.py:172-211).
 *   x_int = clamp(rint(x / scale) + zero_point, int_min, int_max);  y = scale * (x_int - zero_point)
 * IEEE division, round-half-to-even, NaN propagates.  y must not alias x partially (y == x is
 * allowed).  Per-tensor: q.n_params == 1. */
int tq_qdq_f32(const float* x, float* y, int64_t n, tq_qspec q, void* stream);

/* Per-axis variant: x viewed as [outer, C, inner]; parameter c applies to x[:, c, :].
 *   per-embedding / PEG activations (B,T,d), axis=2 -> outer=B*T, C=d, inner=1   (quantizers.py:213-217)
 *   per-channel weights (C_out, ...)              -> outer=1, C=C_out, inner=rest (quantizers.py:219-232) */
int tq_qdq_axis_f32(const float* x, float* y, int64_t outer, int64_t C, int64_t inner,
                    tq_qspec q, void* stream);

/* to_integer_forward (quantizers.py:172-187): integer grid only.  Any of the outputs may be NULL.
 *   x_int_f32 : fp32 integers, exactly what the reference returns
 *   x_ctr_bf16: (x_int - zero_point) as bf16 (exact for n_bits <= 8) -- operand format of
 *               tq_linear_qdq_bf16 */
int tq_quant_int_f32(const float* x, float* x_int_f32, void* x_ctr_bf16,
                     int64_t outer, int64_t C, int64_t inner, tq_qspec q, void* stream);

/* ---- range estimation: min / max ------------------------------------------------------------
 * a6-a8: torch.min / torch.max of CurrentMinMax / AllMinMax / RunningMinMax estimators
 * (range_estimators.py:142-143,159-160,206-207).  One pass over x.  out = {min, max}; NaN
 * propagates like torch.  ws: tq_minmax_workspace_bytes() bytes, zeroed once. */
size_t tq_minmax_workspace_bytes(int64_t C);
int tq_minmax_f32(const float* x, int64_t n, float* out_min_max, void* ws, size_t ws_bytes,
                  void* stream);

/* Per-axis min/max without the reference's transpose copy (range_estimators.py:82-85,115-116,
 * 118-120,129-130,178-181,196-197): x viewed as [outer, C, inner] -> mn[C], mx[C]. */
int tq_minmax_axis_f32(const float* x, int64_t outer, int64_t C, int64_t inner,
                       float* mn, float* mx, void* ws, size_t ws_bytes, void* stream);

/* Per-embedding-group statistics (range_estimators.py:87-112, 183-193): groups are contiguous
 * blocks of C/n_groups dims of the (optionally range-sorted) hidden dims.  `ranges` (NULL = no
 * permutation) is the per-dim range vector of the FP32 pass; the sort is stable (ties by index;
 * the reference's torch.argsort leaves ties implementation-defined).  Replaces the reference's
 * dense CxC permutation-matrix matmul by a gather on the [C] vectors (bit-identical).
 * Returns TQ_EINVAL unless C % n_groups == 0 (reference: AssertionError, :89/:185). */
int tq_group_minmax_f32(const float* mn, const float* mx, int64_t C, int32_t n_groups,
                        const float* ranges, float* mn_out, float* mx_out, void* stream);

/* FP32 "ranges" pass for the permutation (range_estimators.py:68-80): ranges = mx - mn; when
 * `first` == 0 the reference's 0.1*r + 0.9*r re-rounding (lines 78-79) is applied. */
int tq_dim_ranges_f32(const float* mn, const float* mx, int64_t C, int32_t first, float* ranges,
                      void* stream);

/* Estimator state update on the device (no host sync):
 *   mode 0 current_minmax : cur = new                              (range_estimators.py:109-116,142-143)
 *   mode 1 running_minmax : cur = first ? new : (1-m)*new + m*cur  (range_estimators.py:205-214)
 *   mode 2 allminmax      : cur = first ? new : min/max(cur, new)  (range_estimators.py:162-167)
 * `momentum` is the python double; (1 - momentum) and momentum are rounded to fp32 like torch does. */
int tq_range_update_f32(const float* new_min, const float* new_max, float* cur_min, float* cur_max,
                        int64_t k, int32_t mode, double momentum, int32_t first, void* stream);

/* ---- set_quant_range on the device ------------------------------------------------------------
 * a4: AsymmetricUniformQuantizer.set_quant_range (quantizers.py:234-282):
 *   x_min = min(x_min, 0); x_max = max(x_max, eps); delta = (x_max - x_min) / (2^n - 1);
 *   zero_float = -x_min / delta; log-domain: delta = log(delta). */
int tq_set_range_asym_f32(const float* x_min, const float* x_max, int64_t k, int32_t n_bits,
                          float eps, int32_t log_domain, float* delta, float* zero_float,
                          void* stream);
/* a5: SymmetricUniformQuantizer.set_quant_range (quantizers.py:334-344):
 *   signed = any(min(x_min,0) < 0); delta = max(|x_min|, x_max) / int_max(signed). */
int tq_set_range_sym_f32(const float* x_min, const float* x_max, int64_t k, int32_t n_bits,
                         float eps, int32_t log_domain, float* delta, uint8_t* is_signed,
                         void* stream);

/* Calibration-time fused GEMM, second half (reference quantization_manager.py:99-106 after hijacker.py:98-116):
 * tq_linear_qdq_bf16(..., tile_minmax) reduced min / max of its own output into tile_minmax (two ordered-int words,
 * zero-initialised by the caller once); this single launch decodes them, applies the estimator update in place on
 * cur_min / cur_max (mode 0 current, 1 running EMA with `momentum`, 2 all-time min/max; `first`: initialise) and sets the
 * per-tensor quantizer range (asymmetric: delta + zero_float; symmetric: delta + is_signed) -- no min/max pass over
 * the tensor, no separate range_update / set_range launches.  The two words are reset to zero. */
int tq_calib_finalize_f32(void* tile_minmax, float* cur_min, float* cur_max, int32_t mode, double momentum,
                          int32_t first, int32_t symmetric, int32_t n_bits, float eps, int32_t log_domain,
                          float* delta, float* zero_float, void* is_signed, void* stream);

/* ---- MSE range estimator --------------------------------------------------------------------
 * a9: MSE_Estimator.loss_fx (range_estimators.py:248-256) for a whole table of candidate
 * quantizers in ONE read of x:   loss_accum[c] += sum_i (x_i - QDQ_c(x_i))^2,  c in [0, n_cand).
 * The candidate table is built by the host exactly like MSE_Estimator.quantize (:287-294) builds
 * its temporary quantizer: cand = {scale, zero_point, int_min, int_max} x n_cand, fp32, laid out
 * as four consecutive arrays of n_cand.  Serves the 1-D grid (:356-376), the 2-D grid (:378-420)
 * and -- with n_cand == 1 -- the golden-section objective (:296-327).
 * Squared errors are fp32, block partials fp32, cross-block accumulation fp64 in a fixed order
 * (deterministic).  ws: tq_mse_workspace_bytes(n_cand) bytes. */
size_t tq_mse_workspace_bytes(int32_t n_cand);
int tq_mse_sse_f32(const float* x, int64_t n, const float* cand, int32_t n_cand,
                   double* loss_accum, void* ws, size_t ws_bytes, void* stream);

/* argmin over the accumulated losses (np.argmin / unravel_index, range_estimators.py:370,406-408:
 * first minimum in C order) and gather of the winning range:
 *   xmin_out[0] = cand_xmin[idx], xmax_out[0] = cand_xmax[idx], idx_out[0] = idx. */
int tq_mse_argmin_f64(const double* loss, int32_t n_cand, const float* cand_xmin,
                      const float* cand_xmax, float* xmin_out, float* xmax_out, int32_t* idx_out,
                      void* stream);

/* ---- hijacked nn.Linear ---------------------------------------------------------------------
 * a11: QuantizationHijacker.forward for QuantLinear (hijacker.py:66-116, autoquant_utils.py:16-21):
 *     y = act_quant( act_fn( x @ Wq.T + bias ) )
 * as one TMA-fed tcgen05 GEMM with a fused epilogue.  Operands are the INTEGER grids of the
 * fake-quantized tensors carried in bf16 (exact for |v| <= 256):
 *     a_ctr [M, K]  = x_int - zero_point of the input activation (tq_quant_int_f32 output)
 *     w_ctr [N, K]  = w_int of the weight (symmetric: zero_point 0)
 * fp32 accumulation in TMEM; epilogue: v = acc * (a_scale * w_scale[n]) + bias[n]; act_fn;
 * a_scale / w_scale[n] are resolved from the operand quantizers `a_q` (per-tensor; delta == NULL:
 * scale 1) and `w_q` (w_q_params = 1 per-tensor or N per output channel);
 * then the output quantizer `out_q` (n_params 1 or N, i.e. per-tensor or per-column/PEG), written
 * as fp32 `y` (may be NULL) and/or the bf16 centred integer grid `y_ctr` (may be NULL).
 * If out_q.delta == NULL the epilogue stops after act_fn (FP32Acts / calibration pass) and, when
 * `tile_minmax` != NULL, the per-tensor min/max of the pre-quantization output is reduced into
 * tile_minmax[2] (calibration: range before quantize, quantization_manager.py:99-106).
 * k_split == 3: A holds three bf16 planes [M, 3K] (hi|mid|lo split of an arbitrary fp32 tensor,
 * a_q.delta = NULL) -> fp32-accurate product for inputs that are not on a per-tensor grid.
 * act_fn: 0 none, 1 GELU (erf), 2 ReLU, 3 Tanh.   Requires K % 64 == 0, N % 8 == 0,
 * 16-byte aligned operands; otherwise TQ_EUNSUPPORTED / TQ_EALIGN (callers use a library GEMM then). */
size_t tq_linear_workspace_bytes(int64_t M, int64_t N, int64_t K);
int tq_linear_qdq_bf16(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias,
                       float* y, void* y_ctr_bf16, int64_t M, int64_t N, int64_t K, int32_t k_split,
                       tq_qspec a_q, tq_qspec w_q, int64_t w_q_params,
                       int32_t act_fn, tq_qspec out_q, int64_t out_q_params, float* tile_minmax,
                       void* ws, size_t ws_bytes, void* stream);

/* Same GEMM with the residual branch of the encoder blocks fused into the epilogue (reference
 * models/quantized_bert.py:238-245 attention output, :264-277 FFN output):
 *     g = dequant(out_q(x @ Wq.T + bias));  y = out2_q(g + res_scale * res_ctr)
 * res_ctr [M, N] is the centred bf16 grid of the residual input (the block input), res_q its
 * per-tensor quantizer, out2_q the residual-sum quantizer (1 or N parameters).  y / y_ctr as above. */
int tq_linear_res_qdq_bf16(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias,
                           float* y, void* y_ctr_bf16, int64_t M, int64_t N, int64_t K,
                           tq_qspec a_q, tq_qspec w_q, int64_t w_q_params,
                           tq_qspec out_q, int64_t out_q_params, const void* res_ctr_bf16,
                           tq_qspec res_q, tq_qspec out2_q, int64_t out2_q_params, void* stream);

/* Residual block with the LayerNorm fused into the GEMM epilogue (reference
 * models/quantized_bert.py:238-245 / 264-277: dense -> QDQ -> + input -> QDQ -> QuantLayerNorm,
 * autoquant_utils.py:55-66):
 *     y = out2_q( dequant(out_q(x @ Wq.T + bias)) + res_scale * res_ctr )
 *     z = ln_q( LayerNorm(y; ln_gamma_q, ln_beta, ln_eps) )
 * Only z leaves the chip (z fp32 and / or z_ctr bf16 centred grid).  The CTAs that cover one 128-row
 * panel form a thread-block cluster and exchange the row statistics through distributed shared
 * memory (exact integer sums of the centred integers -> mean / variance in fp64, rounded once: independent of
 * the tiling and of the summation order).  out_q, out2_q and ln_q are per-tensor; the weight
 * quantizer may be per-channel.  N must be a multiple of 128, 192 or 256 with at most 8 tiles per
 * row panel, else TQ_EUNSUPPORTED (callers run tq_linear_res_qdq_bf16 + tq_ln_qdq_bf16). */
int tq_linear_res_ln_qdq_bf16(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias,
                              float* z, void* z_ctr_bf16, int64_t M, int64_t N, int64_t K,
                              tq_qspec a_q, tq_qspec w_q, int64_t w_q_params, tq_qspec out_q,
                              const void* res_ctr_bf16, tq_qspec res_q, tq_qspec out2_q,
                              const float* ln_gamma_q, const float* ln_beta, float ln_eps,
                              tq_qspec ln_q, void* stream);

/* ---- 8-bit integer operand mode ---------------------------------------------------------------
 * The same hijacked-linear / residual-block semantics as tq_linear_qdq_bf16 / tq_linear_res_ln_qdq_bf16
 * with A, W and the residual carried as the quantizers' INTEGER grids x_int in one byte per element
 * (unsigned for asymmetric / unsigned-symmetric quantizers, two's complement for signed-symmetric ones;
 * n_bits <= 8) and multiplied on the int8 tensor cores (tcgen05 kind::i8, int32 accumulators: the
 * integer GEMM is always exact).  The zero point of A is removed with the integer identity
 *     sum_k (a_int - zp) * w[n,k] = acc[n] - zp * w_rowsum[n],      w_rowsum[n] = sum_k w_int[n,k]  (int32, [N]).
 * Outputs: y fp32 and / or y_ctr_bf16 (centred grid, e.g. for tq_attention_*) and / or y_i8 (x_int
 * bytes: the A operand / residual of the next block).  K % 128 == 0, N % 16 == 0. */
int tq_linear_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias,
                     float* y, void* y_ctr_bf16, void* y_i8, int64_t M, int64_t N, int64_t K,
                     tq_qspec a_q, tq_qspec w_q, int64_t w_q_params, int32_t act_fn,
                     tq_qspec out_q, int64_t out_q_params, void* stream);
/* tq_linear_qdq_bf16 (bf16 centred operands) with the output written as x_int bytes: lets a bf16-operand
 * GEMM feed an 8-bit-operand one (act_fn: 0 none, 1 GELU). */
int tq_linear_qdq_bf16_o8(const void* a_ctr_bf16, const void* w_ctr_bf16, const float* bias, void* y_i8,
                          int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q, int64_t w_q_params,
                          int32_t act_fn, tq_qspec out_q, int64_t out_q_params, void* stream);
/* The engine form of tq_linear_qdq_i8: the weight and output quantizers are given per SEGMENT of output columns
 * (w_q and out_q carry nseg parameter slots; segment j covers columns [j * N / nseg, (j + 1) * N / nseg)) -- one
 * segment for a plain hijacked nn.Linear, three for the fused Q | K | V projection (reference
 * models/quantized_bert.py:135-151: three QuantLinear, each with its own per-tensor quantizers).  Same arithmetic
 * as tq_linear_qdq_i8 (bit-identical outputs), leaner kernel: a parameter warp resolves the quantizers once and
 * streams {bias, zero-point correction} per tile, the epilogue has no per-tile set-up and no run-time format
 * switches.  Exactly one of y_ctr_bf16 / y_i8; act_fn 0 (none), 1 (GELU) or 2 (ReLU); ldc: output row stride in
 * elements (0: N; > N lets several GEMMs fill the column blocks of one buffer); K % 128 == 0, (N / nseg) a multiple
 * of 128, 192 or 256, outputs 32-byte aligned; else TQ_EUNSUPPORTED / TQ_EALIGN (callers use tq_linear_qdq_i8). */
int tq_linear_seg_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias,
                         void* y_ctr_bf16, void* y_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q,
                         tq_qspec w_q, tq_qspec out_q, int32_t nseg, int32_t act_fn, int64_t ldc, void* stream);
/* Hijacked nn.Linear followed by MobileBERT's QuantNoNorm (reference models/quantized_mobilebert.py:58-72: the
 * elementwise affine y * weight + bias with fake-quantized parameters and a quantized output), optionally with the
 * quantized residual sum in between (self-output / FFN-output / output-bottleneck blocks, :273-311, 327-404):
 *     k = out_q(x @ Wq.T + bias)                                    [res_i8 == NULL]
 *     k = out2_q( dequant(out_q(x @ Wq.T + bias)) + dequant(res) )  [res_i8 != NULL]
 *     z = nn_q( dequant(k) * nn_weight_q + nn_bias_q )              -> x_int bytes
 * nn_weight_q / nn_bias_q: the NoNorm parameters AFTER their (shared) weight quantizer, fp32 [N].  All quantizers
 * per-tensor.  One pass in the GEMM epilogue: the two (three) intermediate tensors never leave the SM.
 * ldc: output row stride in elements (0: N).  K % 128 == 0, N % 128 == 0. */
int tq_linear_nonorm_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias, void* z_i8,
                            int64_t M, int64_t N, int64_t K, tq_qspec a_q, tq_qspec w_q, tq_qspec out_q,
                            const void* res_i8, tq_qspec res_q, tq_qspec out2_q, const float* nn_weight_q,
                            const float* nn_bias_q, tq_qspec nn_q, int64_t ldc, void* stream);
/* Per-embedding-group (PEG) activations through the int8 pipeline (BASELINE config 3; reference
 * utils/per_embd_quant_utils.py:54-68, quantization/range_estimators.py:82-112, quantizers.py:213-217).
 * A operand: x_int bytes of a tensor whose (scale, zero point) change per group of K / a_groups CONTRACTION columns
 * (a_q carries a_groups parameter slots; groups are contiguous; K / a_groups a multiple of 128).  The k-loop runs
 * group by group into ping-pong TMEM accumulators and the epilogue keeps  sum_g s_a[g] s_w (acc_g - zp_g * rowsum_g)
 * in registers, so every integer partial sum is exact.  w_grp_rowsum: int32 [a_groups][N], sum of w_int over the K
 * columns of each group.  Output side: tiles are 128 columns wide and every quantizer is constant per SEGMENT of
 * seg_width columns (a multiple of 128; N / seg_width segments): w_q / out_q [/ res_q / out2_q / ln_q] carry 1 or
 * N / seg_width slots -- a per-embedding-group OUTPUT quantizer with groups of >= 128 columns is a per-segment one.
 * tq_linear_peg_res_ln_qdq_i8 is the residual + LayerNorm block (cluster of N / 128 <= 8 CTAs; the row statistics
 * combine every member's exact integer sums with its own output scale in fp64).  a_groups = 1 gives per-tensor A
 * with per-group outputs (the FFN-out block).  Exactly one of y_ctr_bf16 / y_i8 for the plain form. */
int tq_linear_peg_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_grp_rowsum, const float* bias,
                         void* y_ctr_bf16, void* y_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q,
                         int32_t a_groups, tq_qspec w_q, int32_t w_params, tq_qspec out_q, int32_t out_params,
                         int64_t seg_width, int32_t act_fn, void* stream);
int tq_linear_peg_res_ln_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_grp_rowsum, const float* bias,
                                void* z_ctr_bf16, void* z_i8, int64_t M, int64_t N, int64_t K, tq_qspec a_q,
                                int32_t a_groups, tq_qspec w_q, int32_t w_params, tq_qspec out_q,
                                int32_t out_params, const void* res_i8, tq_qspec res_q, int32_t res_params,
                                tq_qspec out2_q, int32_t out2_params, const float* ln_gamma_q,
                                const float* ln_beta, float ln_eps, tq_qspec ln_q, int32_t ln_params,
                                int64_t seg_width, void* stream);
/* Classification head in one launch (reference models/quantized_bert.py:525-560 -- QuantizedBertPooler on the first token,
 * then the classifier QuantLinear): row b of the input = x_i8 + b * row_stride (x_int bytes, hidden size D, D % 128 == 0,
 * D <= 1024); pooler [D, D] int8 + tanh + QDQ(pool_q), classifier [>= L, D] int8 + QDQ(cls_q); logits [B, ldl] receive the
 * DEQUANTIZED classifier outputs.  a_q / pool_q must be unsigned grids (zero_float != NULL).  Same arithmetic as two
 * tq_linear_qdq_i8 calls (activation 3, then 0): bit-identical. */
int tq_head_qdq_i8(const void* x_i8, int64_t row_stride, int32_t B, int32_t D, int32_t L, const void* wp_i8, const int32_t* wp_rowsum,
                   const float* bp, tq_qspec a_q, tq_qspec wp_q, tq_qspec pool_q, const void* wc_i8, const int32_t* wc_rowsum,
                   const float* bc, tq_qspec wc_q, tq_qspec cls_q, float* logits, int64_t ldl, void* stream);

/* The ENCODER CHAIN: a list of stages executed by ONE launch (reference models/quantized_bert.py:135-291, all layers).
 * A cluster of N_hidden / 192 CTAs carries one 128-row panel (= one sequence of 128 tokens) through every stage: the A
 * operand of stage i + 1 is what the same cluster wrote in stage i, so only cluster barriers separate the stages -- the
 * whole encoder (Q | K | V of layer 0, then per layer: attention, attention-output + residual + LayerNorm, FFN-in + GELU,
 * FFN-out + residual + LayerNorm, the next layer's Q | K | V) is one kernel.  Same arithmetic as the single kernels
 * (tq_linear_seg_qdq_i8 / tq_linear_res_ln_qdq_i8 / tq_attention_qdq_i8): bit-identical outputs.
 *   kind 0  int8 GEMM, nseg segments, out = bf16 centred grid       kind 1  int8 GEMM + GELU, out = x_int bytes
 *   kind 2  int8 GEMM + residual + LayerNorm, out = x_int bytes (N / 192 <= 8 CTAs per cluster; res_i8, res_q, out2_q, ln_*)
 *   kind 3  attention: a_i8 = the Q | K | V buffer [M, 3 N] bf16 centred grids, N = hidden size, K = heads (N == 64 K,
 *           K divisible by the cluster size, M % 128 == 0), bias = additive mask [M / 128, 128] or NULL, out = context
 *           x_int bytes [M, N]; quantizers: a_q / w_q / res_q = Q / K / V projections, out2_q = scores, ln_q =
 *           probabilities, out_q = context (all per-tensor)
 * At least one kind-2 stage (it fixes the cluster size); (N / tile) divisible by the cluster size in every GEMM stage;
 * K % 128 == 0; else TQ_EUNSUPPORTED and callers launch the stages one by one.  A plan holds the stage descriptors
 * (tensor maps, pointers, quantizer specs) in device memory: create once for fixed buffers, run per forward. */
typedef struct tq_chain_stage {
    const void*    a_i8;        /* [M, K] x_int bytes (kind 3: [M, 3 N] bf16) */
    const void*    w_i8;        /* [N, K] */
    const int32_t* w_rowsum;    /* [N] */
    const float*   bias;        /* [N] or NULL (kind 3: the mask) */
    void*          out;         /* [M, N] bf16 centred grid (kind 0) or x_int bytes (kinds 1, 2, 3) */
    int64_t        N, K;
    tq_qspec       a_q, w_q, out_q;   /* w_q / out_q: nseg slots */
    int32_t        nseg;
    int32_t        kind;
    const void*    res_i8;      /* kind 2: residual x_int bytes [M, N] */
    tq_qspec       res_q, out2_q, ln_q;
    const float*   ln_gamma_q;
    const float*   ln_beta;
    float          ln_eps;
} tq_chain_stage;
int tq_chain_plan_create(const tq_chain_stage* stages, int32_t n_stages, int64_t M, void** plan);
int tq_chain_plan_run(const void* plan, void* stream);
int tq_chain_plan_destroy(void* plan);
int tq_linear_res_ln_qdq_i8(const void* a_i8, const void* w_i8, const int32_t* w_rowsum, const float* bias,
                            float* z, void* z_ctr_bf16, void* z_i8, int64_t M, int64_t N, int64_t K,
                            tq_qspec a_q, tq_qspec w_q, int64_t w_q_params, tq_qspec out_q,
                            const void* res_i8, tq_qspec res_q, tq_qspec out2_q,
                            const float* ln_gamma_q, const float* ln_beta, float ln_eps,
                            tq_qspec ln_q, void* stream);

/* ---- fused encoder blocks --------------------------------------------
 * All tensors are centred integer grids in bf16 (x_int - zero_point; dequantized value = scale*ctr).
 *
 * Self-attention core of one layer (reference models/quantized_bert.py:153-213):
 *   scores = Q K^T -> s_q QDQ -> / sqrt(head_dim) + mask -> softmax -> p_q QDQ -> P V -> c_q QDQ
 * qkv_ctr [B*T, 3*H*head_dim] (Q | K | V column blocks, head h at columns h*head_dim..), output
 * c_ctr [B*T, H*head_dim].  One CTA per (batch, head), both products on tcgen05 with exact integer
 * operands; scores / probabilities stay in TMEM / shared memory.  q_q, k_q, v_q: per-tensor
 * quantizers of the three projections (scales); mask [B, T] additive fp32 or NULL.
 * Supported: T == 128, head_dim == 64 (else TQ_EUNSUPPORTED: callers use the unfused path). */
int tq_attention_qdq_bf16(const void* qkv_ctr_bf16, void* c_ctr_bf16, int32_t B, int32_t T, int32_t H,
                          int32_t head_dim, tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, tq_qspec s_q,
                          tq_qspec p_q, tq_qspec c_q, const float* mask, void* stream);

/* tq_attention_qdq_bf16 with the context written as x_int bytes (A operand of tq_linear_*_i8). */
int tq_attention_qdq_i8(const void* qkv_ctr_bf16, void* c_i8, int32_t B, int32_t T, int32_t H,
                        int32_t head_dim, tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, tq_qspec s_q,
                        tq_qspec p_q, tq_qspec c_q, const float* mask, void* stream);
/* ... with per-embedding-group quantizers on Q / K / V and on the context whose groups hold whole heads: q_q, k_q,
 * v_q carry qkv_params slots and c_q c_params slots (1 or a divisor of H); head h uses slot h / (H / params). */
/* ... for heads narrower than the kernel's 64: every head occupies a 64-column slot of qkv whose upper columns are
 * ZERO (the projections are run with zero-padded weight rows, so the padding is exact), the scores are divided by
 * sqrt(true_head_dim) -- a true division unless that is a power of two -- and the padded context columns come out
 * as the zero point.  MobileBERT: 4 heads x 32 (reference models/quantized_mobilebert.py:166-270). */
int tq_attention_pad_qdq_i8(const void* qkv_ctr_bf16, void* c_i8, int32_t B, int32_t T, int32_t H,
                            int32_t head_dim, int32_t true_head_dim, tq_qspec q_q, tq_qspec k_q, tq_qspec v_q,
                            tq_qspec s_q, tq_qspec p_q, tq_qspec c_q, const float* mask, void* stream);
int tq_attention_peg_qdq_i8(const void* qkv_ctr_bf16, void* c_i8, int32_t B, int32_t T, int32_t H,
                            int32_t head_dim, tq_qspec q_q, tq_qspec k_q, tq_qspec v_q, int32_t qkv_params,
                            tq_qspec s_q, tq_qspec p_q, tq_qspec c_q, int32_t c_params, const float* mask,
                            void* stream);

/* QuantLayerNorm over a quantized input (reference autoquant_utils.py:55-66 applied to the output of
 * a residual quantizer): x = in_scale * x_ctr; y = LayerNorm(x; gamma_q, beta, eps); out_q QDQ.
 * gamma_q is the fake-quantized LayerNorm weight.  in_q / out_q: 1 or D parameters.  out_f32 optional.
 * One warp per row; requires D % 256 == 0, D <= 1024. */
int tq_ln_qdq_bf16(const void* x_ctr_bf16, tq_qspec in_q, int64_t in_q_params, const float* gamma_q,
                   const float* beta, float eps, tq_qspec out_q, int64_t out_q_params,
                   void* out_ctr_bf16, float* out_f32, int64_t M, int32_t D, void* stream);

/* Embedding block (reference models/quantized_bert.py:59-88): e_tok QDQ(word[id] + type[tt]) ->
 * e_pos QDQ(. + pos[p]) -> LayerNorm -> out_q QDQ.  Tables are the fake-quantized embedding weights
 * (fp32).  type_ids / pos_ids may be NULL (all zero / row % T). */
int tq_embed_ln_qdq_bf16(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int64_t T,
                         const float* word_q, const float* type_q, const float* pos_q,
                         tq_qspec e_tok, int64_t e_tok_params, tq_qspec e_pos, int64_t e_pos_params,
                         const float* gamma_q, const float* beta, float eps, tq_qspec out_q,
                         int64_t out_q_params, void* out_ctr_bf16, float* out_f32, int64_t M, int32_t D,
                         void* stream);

/* hi|mid|lo bf16 split of an fp32 tensor [M, K] -> [M, 3K] (x == hi + mid + lo to 2^-24 rel.). */
int tq_split3_bf16(const float* x, void* out_bf16, int64_t M, int64_t K, void* stream);

/* tq_embed_ln_qdq_bf16 with the output written as x_int bytes (8-bit operand mode). */
int tq_embed_ln_qdq_i8(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int64_t T,
                       const float* word_q, const float* type_q, const float* pos_q, tq_qspec e_tok,
                       int64_t e_tok_params, tq_qspec e_pos, int64_t e_pos_params, const float* gamma_q,
                       const float* beta, float eps, tq_qspec out_q, int64_t out_q_params,
                       void* out_i8, int64_t M, int32_t D, void* stream);

/* ---- training-time quantizer path -----------------------------------
 * Straight-through backward of the fake quantizer with learnable ranges: autograd of
 * AsymmetricUniformQuantizer.forward / SymmetricUniformQuantizer.forward (quantizers.py:172-211,
 * round_ste :12-20, scale :142-147, zero_point :149-153; make_range_trainable :284-288, 346-349)
 * for the upstream gradient grad_y, x viewed [outer, C, inner] as in tq_qdq_axis_f32:
 *     u = rint(x / s) + zp;  in = (int_min <= u <= int_max);  h = grad_y * s
 *     grad_x          = in ? h / s : 0
 *     grad_scale[c]   = sum grad_y * (clamp(u) - zp)  -  sum (in ? h : 0) * ((x / s) / s)
 *     grad_zp[c]      = sum (in ? 0 : -h)
 *     grad_delta[c]   = delta >= eps ? grad_scale : 0          ('log' domain: grad_scale * s)
 *     grad_zero_float = int_min <= rint(zero_float) <= int_max ? grad_zp : 0
 * One pass: 12 B / element (the reference's autograd graph makes ~14 passes).  Any of grad_x,
 * grad_delta[n_params], grad_zero_float[n_params] may be NULL; grad_zero_float is ignored for a
 * symmetric quantizer.  Sums: fp32 per thread, fp64 across threads / CTAs in a fixed order
 * (deterministic).  ws: tq_qdq_bwd_workspace_bytes() bytes, 16-byte aligned, zeroed once; one buffer
 * (sized for the largest call) may serve all calls of a stream. */
size_t tq_qdq_bwd_workspace_bytes(int64_t outer, int64_t C, int64_t inner);
int tq_qdq_bwd_f32(const float* x, const float* grad_y, float* grad_x, float* grad_delta,
                   float* grad_zero_float, int64_t outer, int64_t C, int64_t inner, tq_qspec q,
                   void* ws, size_t ws_bytes, void* stream);

/* AdaRound soft-rounding quantizer, AdaRoundQuantizer.to_integer_forward in a relaxation mode
 * (quantization/adaround/quantizer.py:46-92).  mode: 0 learned_sigmoid, 1 learned_hard_sigmoid
 * (zeta 1.1, gamma -0.1), 2 sigmoid_temp_decay (needs temperature > 0).  w viewed [outer, C, inner].
 *   init : alpha such that the soft target equals the rounding rest x/s - floor(x/s)       (:54-71)
 *   fwd  : x_int = clamp(floor(w / s) + (soft ? h(alpha) : alpha >= 0) + zp, int_min, int_max);
 *          y = s * (x_int - zp).  y and / or x_int may be NULL                               (:73-82)
 *   bwd  : grad_alpha = grad_y * s * [int_min <= u <= int_max] * h'(alpha)   (soft targets) */
int tq_adaround_init_alpha_f32(const float* w, float* alpha, int64_t outer, int64_t C, int64_t inner,
                               tq_qspec q, int32_t mode, float temperature, void* stream);
int tq_adaround_fwd_f32(const float* w, const float* alpha, float* y, float* x_int, int64_t outer,
                        int64_t C, int64_t inner, tq_qspec q, int32_t mode, int32_t soft_targets,
                        float temperature, void* stream);
int tq_adaround_bwd_f32(const float* w, const float* alpha, const float* grad_y, float* grad_alpha,
                        int64_t outer, int64_t C, int64_t inner, tq_qspec q, int32_t mode,
                        float temperature, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TQ_B200_H */
