/*
 * dolomite_b200.h -- C ABI of the B200-native (sm_100a) hot path of dolomite-engine.
 *
 * Scope: the data-parallel GPTDolomite / MoEDolomite training step.  The reference
 * (ibm-granite/dolomite-engine @ 2024_08_07) is pure Python and reaches its GPU kernels through
 * torch / flash-attn / scattermoe; it has no FFI of its own.  Every entry point below therefore cites
 * the reference *call site* (file:line under dolomite_engine/) whose arithmetic it replaces.  The Python
 * host side (dolomite_engine_b200/) binds these with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (DOLO_OK) or a negative error code; dolomite_b200_last_error() returns a
 *     thread-local human readable message for the last failure on the calling thread.
 *   - all pointers are DEVICE pointers unless the name ends in _host.  The callee never allocates, frees
 *     or retains device memory and never synchronises the stream.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - activations/weights are bf16 (uint16 storage), statistics / master weights / gradients-of-weights fp32.
 *   - row-major everywhere; `ld*` are leading dimensions in ELEMENTS.
 */
#ifndef DOLOMITE_B200_H
#define DOLOMITE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOLOMITE_B200_ABI_VERSION 1

#define DOLO_OK 0
#define DOLO_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define DOLO_ERR_CUDA (-2)    /* CUDA runtime / driver error     */

/* library management (no reference counterpart): thread-local text of the last error, ABI version, device query */
const char* dolomite_b200_last_error(void);
int dolomite_b200_abi_version(void);
int dolomite_b200_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* process-wide tuning knobs:
 *   "gemm_cta_pair"    0 | 1 (dense GEMMs on 2-CTA clusters with tcgen05.mma.cta_group::2)
 *   "gemm_sm_margin"   SMs the persistent GEMM grids leave free for concurrent communication kernels (static schedule only)
 *   "gemm_dynamic"     1 | 0 (default 1: one cluster per output tile; running clusters take over pending ones through cluster launch
 *                      control, so the grid uses every SM that is or becomes free -- no margin needed next to NCCL kernels)
 *   "attn_fwd_split"   1 | 0 | 2 | 3 (split-softmax attention forward -- one CTA per SM, double-buffered scores in TMEM, several
 *                      threads per query row -- for head_dim >= 96 (1, default), never (0), also for head_dim 64 / 80 (2);
 *                      3 = like 2 with FOUR instead of two threads per query row)
 *   "attn_head_fastest" heads per chunk of the attention CTA order (default 8: inside a chunk heads fastest + longest tiles
 *                      first, so that the last wave is short tiles; 0 = tiles fastest, round 1's order)
 *   "gemm_l2_hints"    1 | 0 (long-contraction GEMMs load the streamed operand evict-first and the re-used one evict-last)
 *   "gemm_f32_tma_epilogue" 0 | 1 (fp32 weight gradients through TMA tile store / reduce-add instead of per-thread stores) */
int dolomite_b200_set_option(const char* key, int value);
int dolomite_b200_get_option(const char* key, int* value);

/* ------------------------------------------------------------------------------------------------
 * RMSNorm  -- hf_models/modeling_utils/normalization/rmsnorm/base.py:18-25
 *   y = w * bf16( x32 * rsqrt(mean(x32^2) + eps) ), rstd saved for backward.
 *   bwd: dx (+ optional dx_add, the residual-stream gradient), dw accumulated (+=) into fp32.
 *   workspace for bwd: dolomite_b200_rmsnorm_bwd_workspace_bytes(H) bytes.
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t T, int H, float eps,
                              void* stream);
int64_t dolomite_b200_rmsnorm_bwd_workspace_bytes(int H);
int dolomite_b200_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dx_add,
                              void* dx, float* dw_accum, void* workspace, int64_t T, int H, void* stream);

/* ------------------------------------------------------------------------------------------------
 * RoPE on the packed qkv buffer, in place -- hf_models/modeling_utils/position_embedding/rope.py:104-114,
 *   call sites attention/padding_free.py:38-40; cos/sin gather gpt_dolomite/base.py:289-296.
 *   qkv rows are `n_groups` groups of (q_per_group + 2) head slots of head_dim (layouts of
 *   attention/padding_free.py:79-116: mha = nh groups x [q,k,v]; gqa = nkv groups x [q*g,k,v]; mqa = 1 group).
 *   The first q_per_group+1 slots of each group (queries and the key) are rotated.
 *   cos/sin: bf16 tables [n_positions, head_dim]; position_ids int32 or int64 [T].
 *   inverse != 0 applies the transpose rotation (backward).
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_rope_qk_inplace(void* qkv, int64_t row_stride, int64_t T, int n_groups, int q_per_group,
                                  int head_dim, const void* cos_table, const void* sin_table, const void* position_ids,
                                  int position_ids_is_int64, int64_t n_positions, int inverse, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm -- normalization_function "layernorm" = torch.nn.LayerNorm
 * (hf_models/modeling_utils/normalization/layernorm/__init__.py): fp32 statistics, one rounding to bf16:
 *   y = bf16((x - mean) * rstd * w + b);  b may be null.  mean / rstd (fp32 [T]) are saved for backward.
 * bwd: dx = rstd * (g*w - mean(g*w) - xhat * mean(g*w*xhat)) [+ dx_add];  dw_accum += sum g*xhat;  db_accum += sum g.
 * workspace: dolomite_b200_layernorm_bwd_workspace_bytes(H) bytes, 16-byte aligned.
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t T,
                                int H, float eps, void* stream);
int64_t dolomite_b200_layernorm_bwd_workspace_bytes(int H);
int dolomite_b200_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd,
                                const void* dx_add, void* dx, float* dw_accum, float* db_accum, void* workspace, int64_t T,
                                int H, void* stream);

/* ------------------------------------------------------------------------------------------------
 * tanh-GELU -- activation_function "gelu_pytorch_tanh" (hf_models/modeling_utils/activations/base.py), the non-GLU MLP
 * of gpt_dolomite/mlp.py:45-50.  bwd: dx = dy * gelu'(x); dbias_accum (fp32 [F], may be null) += column sums of the
 * bf16 dx (bias gradient of c_fc).
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_gelu_fwd(const void* x, void* y, int64_t n, void* stream);
int dolomite_b200_gelu_bwd(const void* dy, const void* x, void* dx, float* dbias_accum, int64_t T, int64_t F, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SwiGLU -- hf_models/modeling_utils/activations/glu.py:26-28 with gpt_dolomite/mlp.py:54-55 ordering:
 *   x = [up | gate] (first F columns up, last F gate);  y = up * silu(gate).
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_swiglu_fwd(const void* x, void* y, int64_t T, int64_t F, void* stream);
int dolomite_b200_swiglu_bwd(const void* dy, const void* x, void* dx, int64_t T, int64_t F, void* stream);
/* Same, and additionally accumulates the bias gradient of the linear layer that produced x (autograd of
 * ParameterizedLinear's `+ bias`, modeling_utils/linear.py): dbias_accum[0..2F) += column sums of the bf16 dx written. */
int dolomite_b200_swiglu_bwd_bias(const void* dy, const void* x, void* dx, float* dbias_accum, int64_t T, int64_t F,
                                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embedding -- gpt_dolomite/base.py:351-372 (wte gather, * m_emb); bwd accumulates (+=) into fp32 dwte.
 *   ids int64 [T].  Out-of-range ids are an error on the host side (checked by the caller), the kernel clamps.
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_embedding_fwd(const int64_t* ids, const void* wte, void* out, int64_t T, int H, int64_t V,
                                float scale, void* stream);
int dolomite_b200_embedding_bwd(const int64_t* ids, const void* dout, float* dwte, int64_t T, int H, int64_t V,
                                float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cross entropy (mean over non-ignored tokens), fused forward + backward --
 *   model_wrapper/pretraining.py:124-125 (F.cross_entropy on [T,V]) and gpt_dolomite/main.py:185-200.
 *   logits bf16 [T, ldl]; labels int64 [T]; label == ignore_index contributes nothing.
 *   Writes per-token loss (fp32, 0 for ignored), the scalar mean loss, and overwrites dlogits (may alias
 *   logits) with (softmax - onehot) * grad_scale / n_valid in bf16.
 *   scratch: 2 floats.  logit_scale multiplies logits before the softmax (1/m_width, gpt_dolomite/main.py:155-156).
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_cross_entropy_fwd_bwd(const void* logits, int64_t ldl, const int64_t* labels, void* dlogits,
                                        float* loss_per_token, float* loss_mean, float* scratch, int64_t T, int64_t V,
                                        int64_t ignore_index, float logit_scale, float grad_scale, void* stream);
/* The same computation in three steps, for an LM head that never materialises [T, V] (gpt_dolomite/main.py:172-177 +
 * model_wrapper/pretraining.py:107-127 fused): `_count` leaves the number of labels != ignore_index of the WHOLE batch in
 * scratch[0]; `_rows` handles any chunk of rows (logits of the chunk, its labels, its slice of loss_per_token) and may be
 * called once per chunk; `_mean` reduces loss_per_token [T] to the scalar loss.  A label outside [0, V) that is not
 * ignore_index traps (device-side assert, like torch). */
int dolomite_b200_cross_entropy_count(const int64_t* labels, int64_t T, int64_t ignore_index, float* scratch,
                                      void* stream);
int dolomite_b200_cross_entropy_rows(const void* logits, int64_t ldl, const int64_t* labels, void* dlogits,
                                     float* loss_per_token, const float* scratch, int64_t T, int64_t V,
                                     int64_t ignore_index, float logit_scale, float grad_scale, void* stream);
int dolomite_b200_cross_entropy_mean(const float* loss_per_token, int64_t T, const float* scratch, float* loss_mean,
                                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Column sum (bias gradient of nn.Linear, autograd of linear.py:5-25):  out[n] += scale * sum_t x[t, n]
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_colsum_accum(const void* x, int64_t ldx, float* out, int64_t T, int64_t N, float scale,
                               void* stream);
/* x[i] *= scale[0]  (bf16 in place; scale is a DEVICE scalar: the upstream gradient autograd hands to the loss when the
 * caller does anything but `loss.backward()`, train_utils.py:61-90) */
int dolomite_b200_scale_bf16_by_device_scalar(void* x, int64_t n, const float* scale, void* stream);

/* out = a + alpha * b  (bf16; residual adds of gpt_dolomite/layer.py:70-85), a/b/out may alias */
int dolomite_b200_add_scaled(const void* a, const void* b, void* out, float alpha, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training-mode dropout on flat bf16 activations -- nn.Dropout at gpt_dolomite/base.py:138 (after the embeddings),
 * attention/base.py:92 + padding_free.py:75 (after the attention c_proj), gpt_dolomite/mlp.py:43-49 (after the MLP c_proj),
 * moe_dolomite/moe/base.py:106-120 -- fused with the `* m_residual` / `* m_emb` and `+ residual` that follow it
 * (gpt_dolomite/layer.py:73-86, base.py:368-371), each with the reference's own bf16 rounding:
 *   fwd: out = [residual +] bf16(bf16(x * s) * post_mul)      s = 1 / (1 - p) on kept elements, 0 on dropped ones
 *   bwd: dx  = bf16(bf16(dy * pre_mul) * s)
 * The mask is a counter-based hash of (element index, key0, key1): backward and re-computed (checkpointed) blocks regenerate
 * it from the same keys.  p in [0, 1); n % 8 == 0; residual may be NULL; out may alias x / residual, dx may alias dy.
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_dropout_fwd(const void* x, const void* residual, void* out, int64_t n, float p, float post_mul,
                              uint32_t key0, uint32_t key1, void* stream);
int dolomite_b200_dropout_bwd(const void* dy, void* dx, int64_t n, float p, float pre_mul, uint32_t key0, uint32_t key1,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer-side flat-shard kernels (train_utils.py:99-106: clip_grad_norm_ + AdamW step).
 *   sumsq: out[0] += sum(g^2)  (fp32 grads).   clip coef: coef = min(1, max_norm / (sqrt(sumsq) + 1e-6)).
 *   adamw: torch.optim.AdamW semantics on fp32 master shard; also emits the bf16 copy that the next
 *   all-gather ships.  `clip_coef` is a device pointer (nullable -> 1).  step >= 1.
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_sumsq_accum(const float* g, int64_t n, float* out, void* stream);
int dolomite_b200_clip_coef(const float* sumsq, float max_norm, float* coef_out, float* norm_out, void* stream);
int dolomite_b200_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step,
                             const float* clip_coef, void* stream);
/* fp32 -> bf16 cast of a flat shard (FSDP MixedPrecision param_dtype = bf16, distributed/__init__.py:34-44);
 * bf16 reduce-scatter output -> fp32 shard gradient accumulate (reduce_dtype = bf16, same table) */
int dolomite_b200_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int dolomite_b200_accum_bf16_into_f32(const void* src, float* dst, float scale, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Single-query attention over a KV cache (decoding with `past_key_values`: attention/sdpa.py:11-83, attention/flash.py:16-140;
 * model_wrapper/base.py:110-136 `generate`).  One new token per sequence:
 *   qkv      bf16 [batch, row_stride]: the packed c_attn output of the new tokens (RoPE applied); only the q slots are read
 *   k_cache / v_cache  bf16 [batch, L_max, n_groups * head_dim]: keys / values by position (the new token already appended)
 *   lens     int32 [batch]: valid positions per sequence INCLUDING the new token
 *   out      bf16 [batch, n_heads * head_dim]
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_attn_decode(const void* qkv, int64_t row_stride, const void* k_cache, const void* v_cache, const int32_t* lens,
                              void* out, int batch, int64_t L_max, int n_groups, int q_per_group, int head_dim,
                              float softmax_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * bf16 GEMM on tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue), replacing the cuBLAS
 * calls behind nn.Linear (linear.py:5-25; call sites attention/base.py:100, padding_free.py:74,
 * gpt_dolomite/mlp.py:46-48, gpt_dolomite/main.py:172-177) and their autograd (dgrad / wgrad).
 *
 *   D[M,N] = alpha * (sum_k A[m,k] * B[n,k] + bias[n]) + beta * C[m,n]
 *
 *   A is logical [M,K]: a_mn_major == 0 -> stored row-major [M,K] (ld = lda);  1 -> stored [K,M] (ld = lda).
 *   B is logical [N,K]: b_mn_major == 0 -> stored row-major [N,K] (ld = ldb);  1 -> stored [K,N] (ld = ldb).
 *   D/C: row-major [M,N]; d_is_f32 selects fp32 (else bf16) for BOTH D and C.  C may be NULL (beta ignored)
 *   or alias D.  bias: bf16 [N] or NULL.   K % 8 == 0, lds % 8 == 0, 16-byte aligned bases.
 *   flags: bit0 = epilogue via smem staging + TMA store (bf16 D, no C), else direct vector stores.
 * ------------------------------------------------------------------------------------------------ */
#define DOLO_GEMM_FLAG_TMA_STORE 1
/* bit1: D(fp32) += alpha*A.B^T with split-K + fp32 vector atomics (weight gradients); C must be NULL or alias D with
 * beta == 1, bias NULL.  Summation order over K splits is not deterministic (like FSDP's own reduce order). */
#define DOLO_GEMM_FLAG_SPLITK_ACCUMULATE 2
/* bit2 / bit3: force / forbid the CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles on a 2-CTA cluster);
 * default follows the "gemm_cta_pair" option. */
#define DOLO_GEMM_FLAG_CTA_PAIR 4
#define DOLO_GEMM_FLAG_NO_CTA_PAIR 8
/* bit4 / bit5: fp32 D through per-thread 128-byte row segments (the default) / through shared memory + TMA tile store or
 * reduce-add (measured slower, kept for A/B tests; also the "gemm_f32_tma_epilogue" option) */
#define DOLO_GEMM_FLAG_DIRECT_EPILOGUE 16
#define DOLO_GEMM_FLAG_F32_TMA_EPILOGUE 32
int dolomite_b200_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                            void* D, int64_t ldd, int d_is_f32, const void* C, int64_t ldc, float alpha, float beta,
                            const void* bias, int64_t M, int64_t N, int64_t K, int flags, void* stream);

/* Up to 4 weight gradients in ONE persistent launch (the four nn.Linear of a GPTDolomiteBlock, autograd of linear.py:5-25):
 *   dW_i[M_i, N_i] = alpha_i * dY_i^T X_i  (accumulate[i] == 0: overwrite)   or   dW_i += alpha_i * dY_i^T X_i  (!= 0)
 * dY_i bf16 [K, M_i] (ld_dy), X_i bf16 [K, N_i] (ld_x), dW_i fp32 [M_i, N_i] (ld_dw); K = token rows, common to all. */
int dolomite_b200_gemm_bf16_wgrad_multi(int n_problems, const void* const* dY, const int64_t* ld_dy, const void* const* X,
                                        const int64_t* ld_x, float* const* dW, const int64_t* ld_dw, const int64_t* M,
                                        const int64_t* N, int64_t K, const float* alpha, const int* accumulate,
                                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Grouped GEMM for MoE experts (replaces scattermoe `parallel_linear`, moe_dolomite/moe/scatter.py:38-49, and the
 * per-expert F.linear loop of moe/base.py:12-50).  Token rows are grouped by expert, each segment padded to a
 * multiple of 256 rows (scattermoe `padded_block_indices`); m_tile_group[i] = expert of 128-row tile i (-1: unused).
 *   grouped_m:  D[rows, N] = alpha * A[rows, K] . W[g]^T      W stored [G, N, K] (b_mn_major = 0: expert forward)
 *                                                         or W stored [G, K, N] (b_mn_major = 1: expert dgrad)
 *   grouped_k:  D[g][M, N] = alpha * A_g^T B_g + beta * D[g]  (expert wgrad, fp32): A [K_max, M], B [K_max, N] row-major,
 *               contraction over the rows [group_k_offsets[g], group_k_offsets[g+1]) of expert g.  beta = 0 OVERWRITES: D need not
 *               be initialised, and the slice of an expert with an empty row range is written as zeros; beta != 0 leaves it alone.
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_gemm_bf16_grouped_m(const void* A, int64_t lda, const void* B, int64_t ldb, int b_mn_major, void* D,
                                      int64_t ldd, float alpha, int64_t M_max, int64_t N, int64_t K,
                                      const int32_t* m_tile_group, int num_groups, int flags, void* stream);
/* grouped_m with the gather fused into the operand load (TMA gather4): A is the UNGROUPED [a_rows, K] activation matrix,
 * a_row_index[r] (int32, 16-byte aligned, M_max entries) the source row of grouped row r; W stored [G, N, K]. */
int dolomite_b200_gemm_bf16_grouped_m_gather(const void* A, int64_t lda, int64_t a_rows, const int32_t* a_row_index,
                                             const void* B, int64_t ldb, void* D, int64_t ldd, float alpha, int64_t M_max,
                                             int64_t N, int64_t K, const int32_t* m_tile_group, int num_groups, int flags,
                                             void* stream);
int dolomite_b200_gemm_bf16_grouped_k(const void* A, int64_t lda, const void* B, int64_t ldb, float* D, int64_t ldd,
                                      float alpha, float beta, int64_t M, int64_t N, int64_t K_max,
                                      const int32_t* group_k_offsets, int num_groups, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MoE routing / dispatch (moe_dolomite/moe/base.py:108-181, moe/scatter.py:109-138); all on the device, no host sync.
 *   moe_max_rows: upper bound of padded grouped rows for T tokens (buffer sizing).
 *   moe_route: router logits bf16 [T,E] -> top-k expert ids int32 [T,k] (arg-max order), fp32 softmax weights [T,k],
 *              counts[E] (== bincount, bit exact), offsets_padded[E+1], m_tile_group[max_rows/128], cursors[E] scratch,
 *              row_of_slot[T*k] (grouped row of token-slot), slot_of_row[max_rows] (-1 = padding row),
 *              token_of_row[max_rows] (source token of a grouped row; padding rows name token 0) -- the row index of
 *              dolomite_b200_gemm_bf16_grouped_m_gather.  Segments are padded to 256 rows (CTA-pair super tiles).
 *   moe_gather: X_g[row] = x[slot_of_row[row] / k] or 0.     moe_combine: out = c + alpha * sum_j w_j * Y_g[row_j].
 *   moe_combine_bwd / moe_token_sum / moe_router_bwd: their backward pieces.
 * ------------------------------------------------------------------------------------------------ */
int64_t dolomite_b200_moe_max_rows(int64_t T, int E, int k);
int dolomite_b200_moe_route(const void* router_logits, int64_t T, int E, int k, int32_t* sel_idx, float* sel_w,
                            int32_t* counts, int32_t* offsets_padded, int32_t* m_tile_group, int32_t* cursors,
                            int32_t* row_of_slot, int32_t* slot_of_row, int32_t* token_of_row, void* stream);
int dolomite_b200_moe_gather(const void* x, void* xg, const int32_t* slot_of_row, const int32_t* offsets_padded,
                             int64_t T, int E, int k, int H, void* stream);
int dolomite_b200_moe_combine(const void* yg, const int32_t* row_of_slot, const float* sel_w, const void* c, void* out,
                              int64_t T, int k, int H, float alpha, void* stream);
int dolomite_b200_moe_combine_bwd(const void* dy, const void* yg, const int32_t* slot_of_row,
                                  const int32_t* offsets_padded, const float* sel_w, void* dyg, float* dw, int64_t T,
                                  int E, int k, int H, float alpha, void* stream);
int dolomite_b200_moe_token_sum(const void* dxg, const int32_t* row_of_slot, void* dx, int64_t T, int k, int H,
                                void* stream);
int dolomite_b200_moe_router_bwd(const int32_t* sel_idx, const float* sel_w, const float* dw, void* dlogits, int64_t T,
                                 int E, int k, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Packed var-len causal attention (replaces flash_attn_varlen_func at attention/padding_free.py:51-62).
 *   qkv: packed projection output [T, row_stride] in the slot layout described at rope_qk_inplace.
 *   out: [T, n_heads*head_dim] bf16; lse: fp32 [n_heads, T] (natural log-sum-exp of scale*s).
 *   cu_seqlens int32 [B+1] (same for q and k), causal within each document.
 *   bwd: dqkv has the same layout as qkv (dq, dk, dv written into their slots; bf16).
 *   workspace sizes via the *_workspace_bytes helpers.
 * ------------------------------------------------------------------------------------------------ */
int dolomite_b200_attn_varlen_fwd(const void* qkv, int64_t row_stride, void* out, float* lse,
                                  const int32_t* cu_seqlens, int n_docs, int64_t T, int max_seqlen, int n_groups,
                                  int q_per_group, int head_dim, float softmax_scale, void* stream);
int64_t dolomite_b200_attn_varlen_bwd_workspace_bytes(int64_t T, int n_groups, int q_per_group, int head_dim);
int dolomite_b200_attn_varlen_bwd(const void* dout, const void* qkv, int64_t row_stride, const void* out,
                                  const float* lse, void* dqkv, const int32_t* cu_seqlens, int n_docs, int64_t T,
                                  int max_seqlen, int n_groups, int q_per_group, int head_dim, float softmax_scale,
                                  void* workspace, void* stream);
/* The same with attention-probability dropout (attention/base.py:252 `attn_dropout`; `dropout_p` of flash_attn_varlen_func,
 * attention/padding_free.py:49-59, training mode only): P_ij is kept with probability 1 - dropout_p and scaled by
 * 1 / (1 - dropout_p) after the softmax normaliser was taken over the undropped row.  The mask is a hash of (global query
 * token, global key token, head, key0, key1); backward must be given the keys of its forward.  dropout_p == 0: identical
 * to the functions above. */
int dolomite_b200_attn_varlen_fwd_dropout(const void* qkv, int64_t row_stride, void* out, float* lse,
                                          const int32_t* cu_seqlens, int n_docs, int64_t T, int max_seqlen, int n_groups,
                                          int q_per_group, int head_dim, float softmax_scale, float dropout_p,
                                          uint32_t key0, uint32_t key1, void* stream);
int dolomite_b200_attn_varlen_bwd_dropout(const void* dout, const void* qkv, int64_t row_stride, const void* out,
                                          const float* lse, void* dqkv, const int32_t* cu_seqlens, int n_docs, int64_t T,
                                          int max_seqlen, int n_groups, int q_per_group, int head_dim,
                                          float softmax_scale, float dropout_p, uint32_t key0, uint32_t key1,
                                          void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DOLOMITE_B200_H */
