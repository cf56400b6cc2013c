/* b200_ops.h -- C ABI of libb200.so (transformers_b200/lib/libb200.so), the sm_100a kernels behind the reference's
 * decoder hot path.  huggingface/transformers is pure Python: it has no native FFI of its own for this path, so every
 * entry point below replaces a torch-op sequence of the reference (file:line given per function, relative to
 * /root/reference/src/transformers) and is bound from Python with ctypes (transformers_b200/_lib.py; the binding a
 * reference maintainer would add is shown in INTEGRATION.md).
 *
 * Conventions: all tensors are device pointers owned by the caller (bf16 unless noted); sizes are in elements; strides
 * are in elements; every call is asynchronous on `stream`; return 0 on success, negative errno-style code otherwise
 * (-22 invalid argument, -19 no sm_100 device, -5 CUDA error) with a message available from b200_last_error().
 * Thread-compatible: no global mutable state except one-time function-attribute / driver-entry-point caches.
 */
#ifndef B200_OPS_H
#define B200_OPS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* b200_stream_t; /* == cudaStream_t */

int b200_abi_version(void);
int b200_device_check(void);           /* 0 iff the current device is sm_100 (B200) */
const char* b200_last_error(void);     /* thread-local message of the last failing call */

/* nn.Linear forward / dgrad / wgrad (models/llama/modeling_llama.py:174-176, :254-256, :280, :480):
 * D[M,N] (+)= sum_k A(m,k) B(n,k); a_mn/b_mn = 0: operand stored [M|N, K]; 1: stored [K, M|N].  tcgen05 + TMA. */
int b200_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                   int b_mn, int accumulate, b200_stream_t stream);
/* decode rows, 1 <= M <= 4: y[M,N] = x[M,K] W[N,K]^T as one stream over W (HBM-bound: CUDA cores, one warp per output
 * column); b200_gemm_bf16 dispatches to it for such shapes */
int b200_gemv_bf16(const void* x, const void* W, void* y, int M, int N, int K, int ldx, int ldw, int ldy,
                   b200_stream_t stream);
/* CTA-pair variant (tcgen05 cta_group::2, 256x256 tile per 2-CTA cluster); b200_gemm_bf16 dispatches to it for M > 128 */
int b200_gemm_bf16_2sm(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                       int b_mn, int accumulate, b200_stream_t stream);
/* gate|up projection + gated activation in ONE kernel (LlamaMLP.forward models/llama/modeling_llama.py:174-176): W is the
 * [2I, K] block-interleaved gate / up weight (256-row groups = 128 gate_proj rows + the matching 128 up_proj rows);
 * gu [M, 2I] = A W^T in that interleaved column order (kept for the backward), h [M, I] = bf16(bf16(act(gate)) * up).
 * Bit-identical to b200_gemm_bf16 followed by b200_glu_fwd.  M > 128, I % 128 == 0; gelu: 0 SwiGLU, 1 GeGLU. */
int b200_gemm_glu_bf16(const void* A, const void* W, void* gu, void* h, int M, int I, int K, int lda, int ldw, int ldgu,
                       int ldh, int gelu, b200_stream_t stream);
/* grouped GEMM for mixture-of-experts blocks (MixtralExperts.forward models/mixtral/modeling_mixtral.py:69-93;
 * grouped_mm_experts_forward integrations/moe.py:377-478): rows [offsets[g], offsets[g+1]) of A [M_total, K] x expert g's
 * matrix (B + g*N*K: [N, K]; b_mn: [K, N], the dgrad layout) -> the same rows of C [M_total, N].  `offsets`: int32[groups+1]
 * in DEVICE memory (b200_moe_route's output): one launch for all experts, no host synchronisation.  groups <= 64,
 * N % 64 == 0, b_mn needs K % 64 == 0. */
int b200_gemm_bf16_grouped(const void* A, const void* B, void* C, const int* offsets, int groups, int M_total, int N, int K,
                           int lda, int ldb, int ldc, int b_mn, b200_stream_t stream);
/* 1-CTA variant (128x256 tiles); b200_gemm_bf16 dispatches to it for M <= 128 */
int b200_gemm_bf16_1sm(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, int a_mn,
                       int b_mn, int accumulate, b200_stream_t stream);

/* nn.Embedding gather / scatter-add (models/llama/modeling_llama.py:353,381; scaled variant
 * models/gemma2/modeling_gemma2.py:338-349).  ids int64[T]; err_flag int32[1] set to 1 on out-of-range ids. */
int b200_embedding_fwd(const int64_t* ids, const void* weight, void* out, int T, int H, int V, float scale,
                       int has_scale, int* err_flag, b200_stream_t stream);
int b200_embedding_bwd(const int64_t* ids, const void* dout, void* dweight, int T, int H, int V, int64_t padding_idx,
                       float scale, int has_scale, b200_stream_t stream);

/* LlamaRMSNorm.forward (models/llama/modeling_llama.py:62-67) / Gemma2RMSNorm (models/gemma2/modeling_gemma2.py:55-63),
 * optionally fused with the preceding residual add (modeling_llama.py:317,323): res_out = bf16(x + res_in). */
int b200_rmsnorm_fwd(const void* x, const void* res_in, const void* weight, void* res_out, void* y, float* rstd_out,
                     int T, int H, float eps, int gemma, b200_stream_t stream);
int b200_rmsnorm_bwd_workspace_rows(void); /* workspace = fp32[rows * H] */
int b200_rmsnorm_bwd(const void* dy, const void* x, const void* weight, const float* rstd, void* dx, void* dweight,
                     float* workspace, int T, int H, int gemma, int accumulate_dw, b200_stream_t stream);

/* LlamaRotaryEmbedding.forward (models/llama/modeling_llama.py:113-127): cos / sin bf16 [rows, D] with
 * cos[r, j] = cos[r, j + D/2] = bf16(cosf(inv_freq[j] * float(position_ids[r])) * attention_scaling); inv_freq fp32 [D/2],
 * position_ids int64 [rows = B * S].  Bit-exact against the reference's six torch ops. */
int b200_rope_table(const float* inv_freq, const int64_t* position_ids, void* cos_t, void* sin_t, int rows, int D,
                    float attention_scaling, b200_stream_t stream);

/* apply_rotary_pos_emb (models/llama/modeling_llama.py:130-160), in place on the packed projection buffer
 * qkv[B*S, row_stride]; the first n_rot heads (q then k) are rotated; cos/sin bf16 [cos_batch, S, D]. */
int b200_rope(void* qkv, const void* cos_t, const void* sin_t, int B, int S, int n_rot, int D, int row_stride,
              int cos_batch, int bwd, b200_stream_t stream);

/* LlamaMLP gate: act(gate) * up (models/llama/modeling_llama.py:174-176; activations.py:30-49, :92-103).  `gelu` carries
 * two flags: bit 0 = GeGLU (tanh approximation) instead of SwiGLU; bit 1 = gate / up are block-interleaved in ONE [T, 2I]
 * matrix (256-column groups of 128 gate + 128 up columns: the layout b200_gemm_glu_bf16 writes; pass up = gate + 128). */
int b200_glu_fwd(const void* gate, const void* up, void* out, int T, int I, int ld_gu, int ld_out, int gelu,
                 b200_stream_t stream);
int b200_glu_bwd(const void* dh, const void* gate, const void* up, void* dgate, void* dup, int T, int I, int ld_dh,
                 int ld_gu, int ld_dgu, int gelu, b200_stream_t stream);

/* residual add (models/llama/modeling_llama.py:317,323) */
int b200_add_bf16(const void* a, const void* b, void* out, int64_t n, b200_stream_t stream);

/* Attention core (eager_attention_forward models/llama/modeling_llama.py:191-213; sdpa_attention_forward
 * integrations/sdpa_attention.py:79-170; softcap branch models/gemma2/modeling_gemma2.py:201-208; masks
 * masking_utils.py:76-101).  q [B,Sq,Hq,D], k/v [B,Skv,Hkv,D], out [B,Sq,Hq,D] as strided views given by
 * (batch, row, head) strides; lse fp32 [B,Hq,lse_stride]; kv_start/kv_end optional int32[B] valid-kv ranges. */
int b200_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride, int B, int Sq,
                  int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_rs, int64_t q_hs, int64_t k_bs, int64_t k_rs,
                  int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs, int64_t o_bs, int64_t o_rs, int64_t o_hs,
                  float scale, float softcap, int causal, int window, const int* kv_start, const int* kv_end,
                  b200_stream_t stream);
/* strides: 8 tensors x (batch, row, head) for q, k, v, out, dout, dq, dk, dv; workspace fp32[2*B*Hq*lse_stride] */
int b200_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                  void* dq, void* dk, void* dv, float* workspace, int B, int Sq, int Skv, int Hq, int Hkv, int D,
                  int lse_stride, const int64_t* strides, float scale, float softcap, int causal, int window,
                  const int* kv_start, const int* kv_end, b200_stream_t stream);

/* In-place KV-cache append replacing torch.cat in DynamicLayer.update (cache_utils.py:127-146; Cache.update :1349-1381):
 * writes rows [offset, offset+q_len) of the preallocated [B,H,capacity,D] caches. */
int b200_kv_append(const void* k_new, const void* v_new, void* k_cache, void* v_cache, int B, int H, int q_len, int D,
                   int64_t ks_b, int64_t ks_h, int64_t ks_r, int64_t vs_b, int64_t vs_h, int64_t vs_r, int64_t cs_b,
                   int64_t cs_h, int64_t cs_r, int offset, int capacity, b200_stream_t stream);

/* Mixtral experts path (MixtralExperts.forward models/mixtral/modeling_mixtral.py:69-93; grouped_mm_experts_forward
 * integrations/moe.py:377-478): sort (token, k) pairs by expert, gather rows, [expert GEMMs = b200_gemm_bf16 on row
 * ranges], weighted un-permute.  counts must be zeroed by the caller; offsets has E+1 entries. */
int b200_moe_route(const int64_t* top_k_index, int* counts, int* offsets, int* cursor, int* slot, int* token_of_slot,
                   int T, int topk, int E, b200_stream_t stream);
int b200_moe_gather(const void* x, const int* token_of_slot, void* x_sorted, int nslots, int H, b200_stream_t stream);
int b200_moe_combine(const void* y_sorted, const int* slot, const float* weights, void* out, int T, int topk, int H,
                     b200_stream_t stream);

/* ForCausalLMLoss (loss/loss_utils.py:32-70): shifted labels, fp32 log-sum-exp, mean over valid targets. */
int b200_ce_fwd(const void* logits, const int64_t* labels, float* lse, float* loss_rows, float* loss_out,
                float* denom_out, int B, int S, int V, int ld, int shift, int64_t ignore_index, float num_items,
                b200_stream_t stream);
int b200_ce_bwd(const void* logits, const int64_t* labels, const float* lse, const float* dloss, const float* denom,
                void* dlogits, int B, int S, int V, int ld, int ld_out, int shift, int64_t ignore_index,
                b200_stream_t stream);
/* Same gradient for a vocabulary-sharded lm_head (tp_plan "colwise" on lm_head, models/llama/modeling_llama.py:423, without
 * the gather of "colwise_gather_output"): logits holds this rank's V columns; lse_global is the log-sum-exp over the whole
 * vocabulary (combined across ranks by the caller), target_local the target's column in this shard or a value outside
 * [0, V), row_scale = dloss / denom for valid rows and 0 for ignored ones. */
int b200_ce_bwd_sharded(const void* logits, const int64_t* target_local, const float* lse_global,
                        const float* row_scale, void* dlogits, int T, int V, int ld, int ld_out, b200_stream_t stream);

/* Decode step (q_len == 1 over the KV cache; models/llama/modeling_llama.py:243-281 with past_key_values): the context is
 * split over b200_attn_decode_splits(B, Hkv, Skv) CTAs per (batch, kv head), every K / V byte is read once and shared by the
 * Hq / Hkv query heads; a second kernel merges the partial softmax states.  q / out: [B, 1, Hq, D] (batch, head strides);
 * k / v: [B, Skv, Hkv, D] strided; workspace: B * Hq * nsplit * (D + 2) floats; lse (optional): [B, Hq, lse_stride], entry 0.
 * causal masking is implied (the query is the last position); window / kv_start / kv_end as in b200_attn_fwd. */
int b200_attn_decode_splits(int B, int Hkv, int Skv);
int b200_attn_decode(const void* q, const void* k, const void* v, void* out, float* lse, int lse_stride, float* workspace,
                     int B, int Skv, int Hq, int Hkv, int D, int64_t q_bs, int64_t q_hs, int64_t k_bs, int64_t k_rs,
                     int64_t k_hs, int64_t v_bs, int64_t v_rs, int64_t v_hs, int64_t o_bs, int64_t o_hs, float scale,
                     float softcap, int window, const int* kv_start, const int* kv_end, b200_stream_t stream);

/* ---- optimizer step after the path ---------------------------------------------------------------
 * Trainer clips the global gradient norm (trainer.py:1783-1785, _clip_grad_norm :2538-2542 -> torch.nn.utils.clip_grad_norm_)
 * and calls optimizer.step() (:1788) on torch.optim.AdamW (trainer_optimizer.py:201-208).  Multi-tensor, one launch per
 * param group.  tensor_table: device int64 [n_tensors][6] = {param*, grad*, exp_avg*, exp_avg_sq*, numel, fp32 master* or 0} (bf16 params
 * and grads; moments bf16 or fp32); chunk_map: device int32 [n_chunks][2] = {tensor index, chunk index} with chunks of
 * b200_optim_chunk_elems() elements. */
int b200_optim_chunk_elems(void);
/* torch.optim.AdamW update with step_size = lr / bias_correction1, denom = sqrt(v) / bias_correction2_sqrt + eps; every
 * gradient is multiplied by *grad_scale first when grad_scale != NULL (fused clipping: pass out2 + 1 of b200_grad_norm).
 * state_is_fp32: bit 0 = fp32 moments (else bf16); bit 1 = the table's sixth column points to fp32 master parameters: the
 * update runs on them and the bf16 parameter becomes their rounding. */
int b200_adamw_step(const int64_t* tensor_table, const int32_t* chunk_map, int n_chunks, int state_is_fp32, float lr,
                    float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                    float bias_correction2_sqrt, const float* grad_scale, b200_stream_t stream);
/* out2[0] = L2 norm over all gradients in the table (fp32 partial sums per chunk in partial_ws [n_chunks], combined in
 * double: deterministic), out2[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (1 when max_norm <= 0). */
int b200_grad_norm(const int64_t* tensor_table, const int32_t* chunk_map, int n_chunks, float* partial_ws, float max_norm,
                   float* out2, b200_stream_t stream);
/* grad *= *coef in place (no traffic when *coef == 1). */
int b200_grad_scale(const int64_t* tensor_table, const int32_t* chunk_map, int n_chunks, const float* coef,
                    b200_stream_t stream);

/* ---- tensor-parallel reduction over NVLink peer memory ---------------------------------------------
 * The rowwise layers of the tp_plan produce partial sums that the reference all-reduces (distributed/tensor_parallel.py:
 * 320-328).  Here every rank writes its partial into a peer-mapped buffer; after a barrier each rank pulls the rows it owns
 * from all `world` buffers (peer_ptrs: HOST array of device pointers in rank order, own buffer included), sums them in
 * fp32 in rank order, adds `residual` (may be NULL) and writes bf16: reduce-scatter (+ residual add) as one kernel whose
 * loads are the NVLink transfer.  world in {1, 2, 4, 8}; offset_elems / n_elems multiples of 8. */
/* GEMM + first half of the reduce-scatter in ONE kernel: D = A B^T as b200_gemm_bf16 (CTA-pair tcgen05 kernel), but the
 * epilogue TMA-stores row block r of the output (M / world rows, a multiple of 256) straight into dest_ptrs[r] (HOST array
 * of device pointers: for r != rank a slot in rank r's peer-mapped memory), other ranks' blocks first, so the partial sums
 * cross NVLink tile by tile under the tensor-core work of the following tiles.  Then: barrier, b200_pull_reduce_bf16 over the
 * `world` local slots. */
int b200_gemm_bf16_scatter(const void* A, const void* B, void* const* dest_ptrs, int world, int rank, int M, int N, int K,
                           int lda, int ldb, int ldc, int a_mn, int b_mn, b200_stream_t stream);
int b200_pull_reduce_bf16(const void* const* peer_ptrs, int world, int64_t offset_elems, int64_t n_elems,
                          const void* residual, void* out, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_OPS_H */
