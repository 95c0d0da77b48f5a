/* libvita_b200.so -- C ABI of the B200-native (sm_100a) kernels behind VITA's omni-modal prefill + decode forward.
 *
 * The reference (VITA-MLLM/VITA) is pure Python and has no FFI of its own: every GPU op is reached through
 * torch / transformers / flash-attn / vLLM wheels.  This header is therefore the boundary a maintainer binds with
 * ctypes; each entry point names the reference call site (file:line under the reference tree)
 * whose library kernel it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless stated otherwise;
 *   - bf16 storage (uint16 payload), fp32 accumulation; row-major; weights are [out_features, in_features]
 *     exactly as torch.nn.Linear stores them;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns VITA_OK (0) or a negative error code and records a message retrievable with
 *     vita_last_error(); nothing allocates device memory (workspaces are passed in);
 *   - there is no CPU fallback: on a machine without an sm_100 GPU the launches fail with VITA_ERR_CUDA.
 */
#ifndef VITA_B200_H
#define VITA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITA_OK 0
#define VITA_ERR_INVALID (-1)
#define VITA_ERR_CUDA (-2)

#define VITA_ACT_NONE 0
#define VITA_ACT_GELU 1 /* erf GELU (torch.nn.GELU default) */
#define VITA_ACT_RELU 2

/* ---- library ---------------------------------------------------------------------------------------------- */
int vita_version(void);
const char* vita_last_error(void);
/* number of SMs of the current device (0 if no CUDA device is usable) */
int vita_num_sms(void);
/* counts kernels launched through this library since the last reset (bench.py's gpu_launches) */
int64_t vita_launch_count(int reset);
/* Tunables (each also has an environment default): "pdl", "attn_early", "chain_wait",
 * "tc_prefetch_consts", "tc_wide_route", "tc_l2_ahead", "tc_trigger_lead".
 * Takes effect for launches issued afterwards (a captured CUDA graph keeps what it was captured with). */
int vita_set_option(const char* name, int64_t value);
int64_t vita_get_option(const char* name);

/* Decode-chain completion counters (bs = 1 greedy step).  Between vita_chain_begin and vita_chain_end the chain-capable
 * launches of the calling thread (vita_decode_tc_*, vita_decode_attention, vita_tc_lm_head_argmax) are linked in call
 * order: each bumps its own 64-bit counter when its CTAs are done and polls its predecessor's instead of waiting for
 * the grid dependency (which resolves ~3.5 us after the last CTA has exited).  mem = [1 + n_links] uint64, zeroed once
 * (and again after a failed launch); mem[0] is the step serial that vita_decode_embed(..., serial = mem) bumps once per
 * step.  Purely an accelerator: a counter that does not arrive in time falls back to the hardware dependency. */
int vita_chain_begin(uint64_t* mem, int64_t n_links);
int vita_chain_end(void);

/* ---- dense linear:  C[M,N] = residual + colscale * act(A[M,K] . B[N,K]^T + bias) ----------------------------
 * tcgen05 / TMEM / TMA GEMM.  Replaces nn.Linear -> cuBLAS at
 *   internvit/modeling_intern_vit.py:180 (qkv), :192 (proj, with ls1 + residual :245-247),
 *   :214,216 (fc1+GELU, fc2 with ls2 + residual :249-251); multimodal_projector/builder.py:164-168;
 *   whale/module/layer/attention.py:371-373,381,419; :145-147; component/subsampling.py:34;
 *   component/transformer.py:313; adapter.py:94,104; transformers MixtralAttention q/k/v/o_proj;
 *   lm_head on all rows (vita_mixtral.py:171-173).
 * bias [N] / colscale [N] / residual [M, ldr] may be NULL.  K, lda, ldc, ldr multiples of 8; 16-byte aligned. */
int vita_gemm_bf16(const void* A, int64_t lda, const void* B, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                   const void* bias, int act, const void* colscale, const void* residual, int64_t ldr, void* stream);

/* ---- row kernels ------------------------------------------------------------------------------------------ */
/* transformers MixtralRMSNorm.forward (modeling_mixtral.py:148-153). */
int vita_rmsnorm(const void* x, const void* w, void* y, int64_t rows, int64_t H, float eps, void* stream);
/* torch.nn.LayerNorm + optional activation + output scale: modeling_intern_vit.py:229-230;
 * whale transformer.py:88-89,313-318 (LayerNorm -> ReLU, then x sqrt(d) attention.py:109),:371; adapter.py:98-104. */
int vita_layernorm(const void* x, const void* w, const void* b, void* y, int64_t rows, int64_t H, float eps, int act,
                   float out_scale, void* stream);
/* out[dst_index[i]] = table[src_index[i]] (NULL index = identity, negative = skip): embed_tokens
 * (vita_arch.py:274), placeholder splice (vita_arch.py:277-303; mixtral.py:1116,1126), MoE token gather. */
int vita_row_copy(const void* table, const int32_t* src_index, const int32_t* dst_index, void* out, int64_t n_rows,
                  int64_t H, void* stream);
/* rotate-half RoPE on the q and k heads of a fused qkv activation, in place, and append k/v to the paged cache
 * (modeling_mixtral.py:224-254; vLLM twin mixtral.py:477-501).  cos_sin fp32 [max_pos, 2, D/2]; slot_mapping[tok] =
 * page * page_size + offset (NULL = do not write the cache). */
int vita_rope_kv_write(void* qkv, const int32_t* positions, const int32_t* slot_mapping, const float* cos_sin,
                       void* k_cache, void* v_cache, int64_t n_tok, int64_t n_q_heads, int64_t n_kv_heads,
                       int64_t head_dim, void* stream);

/* The qkv projection with both of the above in its epilogue: qkv_out[M, (n_q + 2 n_kv) * 128] = X[M, K] . W_qkv^T, the
 * q and k heads rotated with positions[row] (on the bf16-rounded projections, bit-identical to vita_gemm_bf16 followed
 * by vita_rope_kv_write), k / v heads also stored at slot_mapping[row] of the paged caches (NULL = no cache write).
 * Replaces q_proj/k_proj/v_proj + apply_rotary_pos_emb + the cache update of MixtralAttention.forward
 * (modeling_mixtral.py:312-340; vLLM twin mixtral.py:470-501). */
int vita_gemm_qkv_rope(const void* X, int64_t ldx, const void* W_qkv, void* qkv_out, int64_t M, int64_t K,
                       int64_t n_q_heads, int64_t n_kv_heads, int64_t head_dim, const int32_t* positions,
                       const int32_t* slot_mapping, const float* cos_sin, void* k_cache, void* v_cache, void* stream);

/* ---- attention -------------------------------------------------------------------------------------------- */
/* softmax(Q K^T * scale [causal] [kv_lens mask]) V, FlashAttention style.  Strides are {batch, token, head} in
 * elements.  Supported (d_qk, d_v): (128,128) Mixtral GQA, (64,64) InternViT, (128,64) Whale rel-pos with the
 * operands prepared by vita_whale_qk_prep.  Replaces flash_attn_varlen_qkvpacked_func
 * (internvit/flash_attention.py:61), whale attention.py:391-415, transformers sdpa (modeling_mixtral.py:269-292).
 * causal: query row i attends keys <= q_pos0 + i (q_pos0 = 0 for a whole sequence; a sequence shard passes the
 * position of its first row). */
int vita_attention_fwd(const void* q, const void* k, const void* v, void* o, const int64_t* q_strides,
                       const int64_t* k_strides, const int64_t* v_strides, const int64_t* o_strides, int64_t B,
                       int64_t n_q_heads, int64_t n_kv_heads, int64_t Sq, int64_t Skv, int64_t d_qk, int64_t d_v,
                       const int32_t* kv_lens, int causal, int64_t q_pos0, float scale, void* stream);
/* single-query paged-KV attention for decode (vLLM paged Attention, mixtral.py:484-501).  workspace must be
 * zero-initialised once and be at least vita_decode_attention_workspace_bytes() large for the largest (B, splits) it
 * is used with; every launch hands it back all-zero, so B and splits may change from call to call.  The context
 * splits of a kv head are merged all-to-all through tagged 64-bit words when splits is 4, 8 or 16 and the grid is
 * resident at once, otherwise by the last CTA to take a ticket ("attn_tagged"). */
int64_t vita_decode_attention_workspace_bytes(int64_t B, int64_t n_kv_heads, int64_t splits);
int vita_decode_attention(const void* q, const void* k_cache, const void* v_cache, const int32_t* block_table,
                          const int32_t* cur_pos, void* out, void* workspace, int64_t B, int64_t n_q_heads,
                          int64_t n_kv_heads, int64_t head_dim, int64_t page_size, int64_t max_pages, int64_t splits,
                          float scale, int64_t q_stride /* elements between batch rows of q; 0 = dense */, void* stream);

/* ---- sparse MoE (prefill) --------------------------------------------------------------------------------- */
/* post_attention_layernorm + MixtralTopKRouter (modeling_mixtral.py:109-116; vLLM FusedMoE renormalize=True,
 * mixtral.py:405-414): writes the normed activations xn, top-2 expert ids and renormalised weights. */
int vita_moe_router(const void* h, const void* norm_w, const void* gate_w, void* xn, int32_t* topk_ids, float* topk_w,
                    int64_t n_tok, int64_t H, int64_t E, float eps, void* stream);
/* stable counting sort of the (token, k) assignments by expert: expert_offsets [E+1], perm_row [n_tok*2],
 * row_token [n_tok*2], row_weight [n_tok*2] (replaces the one_hot / torch.where bookkeeping, modeling_mixtral.py:81-90). */
int vita_moe_align(const int32_t* topk_ids, const float* topk_w, int32_t* expert_offsets, int32_t* perm_row,
                   int32_t* row_token, float* row_weight, int32_t* row_assign /* [n_tok*2] t*2+k per row, may be NULL */,
                   int64_t n_tok, int64_t E, void* stream);
/* grouped expert GEMMs over the permuted rows (modeling_mixtral.py:91-95):
 *   Act[r, :]    = silu(X[r] . Wg_e^T) * (X[r] . Wu_e^T)   with W_gate_up [E, 2I, H] (gate rows first)
 *   Y_perm[r, :] = row_weight[r] * (Act[r] . Wd_e^T)       with W_down [E, H, I] */
int vita_moe_gemm_gate_up_silu(const void* X_perm, const void* W_gate_up, void* Act, const int32_t* expert_offsets,
                               int64_t rows, int64_t num_experts, int64_t H, int64_t I, void* stream);
int vita_moe_gemm_down(const void* Act, const void* W_down, void* Y_perm, const int32_t* expert_offsets,
                       const float* row_weight, int64_t rows, int64_t num_experts, int64_t H, int64_t I, void* stream);
/* h[t] += Y_perm[perm_row[t,0]] + Y_perm[perm_row[t,1]] (index_add_ + residual, modeling_mixtral.py:96,386-389);
 * if next_norm_w != NULL also writes xn_out = RMSNorm(h) for the next layer / final norm. */
/* Fused top-2 router + token permute ("slot" layout): as vita_moe_router, but instead of xn / an offsets pass / a gather
 * pass the CTA of a token claims one row in each of its two experts' slot ranges (expert e owns rows [e * capacity,
 * (e + 1) * capacity) of x_slots; expert_counts[e], zero on entry, counts the claimed rows) and stores the normed
 * activations there; perm_row[t, k] = the claimed row, row_weight[row] = the renormalised routing weight.  topk_ids /
 * topk_w (optional) as vita_moe_router.  The grouped GEMMs *_slots walk the same layout; vita_moe_combine gathers by
 * perm_row.  Row order inside an expert is arrival order; results do not depend on it (rows are independent). */
int vita_moe_route_scatter(const void* h, const void* norm_w, const void* gate_w, void* x_slots, int32_t* expert_counts,
                           int32_t* perm_row, float* row_weight, int32_t* topk_ids, float* topk_w, int64_t n_tok,
                           int64_t H, int64_t E, int64_t capacity, float eps, void* stream);
int vita_moe_gemm_gate_up_silu_slots(const void* X_slots, const void* W_gate_up, void* Act_slots,
                                     const int32_t* expert_counts, int64_t capacity, int64_t rows_hint,
                                     int64_t num_experts, int64_t H, int64_t I, void* stream);
int vita_moe_gemm_down_slots(const void* Act_slots, const void* W_down, void* Y_slots, const int32_t* expert_counts,
                             const float* row_weight, int64_t capacity, int64_t rows_hint, int64_t num_experts, int64_t H,
                             int64_t I, void* stream);
int vita_moe_combine(void* h, const void* y_perm, const int32_t* perm_row, const void* next_norm_w, void* xn_out,
                     int64_t n_tok, int64_t H, float eps, void* stream);

/* ---- expert-parallel MoE over NVLink peer memory (fused compute + collective) -----------------------------
 * vita_moe_gemm_down_ep: the down-projection GEMM whose epilogue stores each (token, k) output row directly into the
 * receive buffer of the rank that owns the token (peer_out[rank] -> [chunk, 2, H] bf16 in symmetric memory).
 * vita_ep_signal / vita_ep_wait: epoch flags in symmetric memory (flags [2][n_ranks] per rank; which = 0 "rows
 * landed", 1 "h/xn gathered").  vita_ep_reduce_norm_gather: the owner sums its two slots per token into the residual
 * stream, applies the next RMSNorm and writes both rows into every rank's h / xn (all-gather by P2P stores). */
int vita_moe_gemm_down_ep(const void* Act, const void* W_down, const int32_t* expert_offsets, const float* row_weight,
                          const int32_t* row_assign, void* const* peer_out, int64_t rows, int64_t num_local_experts,
                          int64_t H, int64_t I, int64_t chunk, void* stream);
int vita_ep_signal(void* const* peer_flags, int64_t which, int64_t n_ranks, int64_t my_rank, int64_t epoch, void* stream);
/* waits for flags[which][src] >= epoch of the sources src < n_wait (n_wait = n_ranks: everybody) */
int vita_ep_wait(const int32_t* my_flags, int64_t which, int64_t n_ranks, int64_t n_wait, int64_t epoch, void* stream);
/* gather = 1: the reduced rows go to every rank's h / xn; gather = 0: to this rank's only (sequence-sharded stream) */
int vita_ep_reduce_norm_gather(const void* rs_buf, const int32_t* my_flags, void* const* peer_h, void* const* peer_xn,
                               const void* next_norm_w, int64_t tok0, int64_t n_owned, int64_t n_ranks, int64_t my_rank,
                               int64_t epoch, int64_t H, float eps, int64_t gather, void* stream);
/* all-gather by P2P stores: up to 4 byte ranges [offset, offset + bytes) of this rank's symmetric buffer
 * (peer_base[my_rank]) are copied to the same offsets of every other rank's buffer.  Offsets are multiples of 16,
 * sizes of 4.  Used for the K/V rows and the routed activations of a sequence shard. */
int vita_ep_push(void* const* peer_base, const int64_t* offsets, const int64_t* bytes, int64_t n_ranges,
                 int64_t n_ranks, int64_t my_rank, void* stream);

/* expert-parallel tail of a decoder layer (NCCL variant): h += y where y is the all-reduced sum of the ranks' partial MoE outputs
 * (each rank ran vita_moe_combine on a zeroed buffer with only its local experts' rows filled); optional next RMSNorm. */
int vita_add_rmsnorm(void* h, const void* y, const void* next_norm_w, void* xn_out, int64_t n_tok, int64_t H, float eps,
                     void* stream);

/* ---- InternViT front / back end --------------------------------------------------------------------------- */
/* im2col for Conv2d(3,1024,k=14,s=14) (modeling_intern_vit.py:80-85,109): out [n_img*(HW/P)^2, k_pad]. */
int vita_vit_im2col(const void* images, void* out, int64_t n_img, int64_t C, int64_t HW, int64_t P, int64_t k_pad,
                    void* stream);
/* cat([cls, patches]) + position_embedding (modeling_intern_vit.py:112-121). */
int vita_vit_assemble(const void* patches, const void* cls, const void* pos, void* out, int64_t n_img, int64_t n_patch,
                      int64_t H, void* stream);
/* drop CLS, x scale, pixel_shuffle(0.5) (internvit_encoder.py:35-53,71-77): [n,1+g*g,C] -> [n,(g/2)^2,4C]. */
int vita_vit_pixel_shuffle(const void* h, void* out, int64_t n_img, int64_t grid, int64_t C, float scale,
                           void* stream);

/* ---- image front end ---------------------------------------------------------------------------------------- */
/* One separable pass of Pillow's 8-bit resampler (what `image.resize(...)` runs inside the reference's
 * dynamic_preprocess, data_utils_video_audio_neg_patch.py:1239,1253; Pillow src/libImaging/Resample.c):
 *   out[o, p, c] = clip8((2^21 + sum_k in[bounds[o][0] + k, p, c] * kk[o][k]) >> 22),  k < bounds[o][1]
 * along `axis` (1 = horizontal pass: [H, W, C] -> [H, out_size, C]; 0 = vertical pass: [H, W, C] -> [out_size, W, C])
 * of an interleaved uint8 image.  kk [out_size, ksize] int32 fixed-point coefficients and bounds [out_size, 2] come
 * from the host (vita_b200/image_frontend.py::resample_tables = Resample.c precompute_coeffs + normalize_coeffs_8bpc).
 * Pillow's order for a 2-D resize: horizontal pass first, then vertical. */
int vita_image_resample_u8(const uint8_t* in, uint8_t* out, int64_t H, int64_t W, int64_t C, int axis,
                           int64_t out_size, const int32_t* kk, const int32_t* bounds, int64_t ksize, void* stream);
/* Tiling (dynamic_preprocess :1241-1250, row-major T x T tiles of a [gj*T, gi*T, 3] image) fused with
 * CLIPImageProcessor's rescale + normalise through a [3, 256] bf16 table (mm_utils.py:30-43):
 * out[tile0 + ty*gi + tx][c][y][x] = lut[c][img[ty*T + y][tx*T + x][c]];  out is [N, 3, T, T] bf16. */
int vita_image_tiles_lut(const uint8_t* img, const void* lut, void* out, int64_t gi, int64_t gj, int64_t T,
                         int64_t tile0, void* stream);

/* ---- Whale audio front end -------------------------------------------------------------------------------- */
/* Kaldi log-mel filterbank = torchaudio.compliance.kaldi.fbank as the reference calls it in
 * audioEncoderProcessor.process (whale/init_model.py:48-56: 25 ms povey window, 10 ms shift, snip_edges, DC removal,
 * pre-emphasis 0.97 with the first sample replicated, 512-point power spectrum, log floored at FLT_EPSILON; dither 0).
 * wave: [n_samples] fp32, already multiplied by 2^15 (init_model.py:47); window: [frame_len]; mel_weights_t:
 * [256, n_mel] (fft bin major) triangular filters; mel_span: [n_mel, 2] = [first, last) non-zero bin of each filter;
 * out: [n_frames, n_mel] fp32 with n_frames = 1 + (n_samples - frame_len) / frame_shift (0 frames: no launch).
 * Output feeds vita_whale_conv1 directly. */
int vita_fbank(const float* wave, int64_t n_samples, const float* window, const float* mel_weights_t,
               const int32_t* mel_span, float* out, int64_t frame_len, int64_t frame_shift, int64_t n_mel,
               float preemph, void* stream);
/* GlobalCMVN (cmvn.py:21-32, mean/istd may be NULL) + Conv2d(1,C,3,2) + ReLU (subsampling.py:28-29); feat fp32
 * [B,T,F]; out channels-last [B,T1,F1,C]. */
int vita_whale_conv1(const float* feat, const float* mean, const float* istd, const void* w, const void* bias,
                     void* out, int64_t B, int64_t T, int64_t F, int64_t C, void* stream);
/* im2col for Conv2d(C,C,3,2) on the channels-last map (subsampling.py:30): out [B*T2*F2, 9*C]. */
int vita_whale_im2col2(const void* in, void* out, int64_t B, int64_t T1, int64_t F1, int64_t C, void* stream);
/* rel-pos attention operands (attention.py:379-398): Q2 = [q+u | q+v], K2 = [k | p], each [B*T, heads, 2*dk]. */
int vita_whale_qk_prep(const void* qkv, const void* p, const void* bias_u, const void* bias_v, void* q2, void* k2,
                       int64_t B, int64_t T, int64_t heads, int64_t dk, void* stream);
/* adapter front (adapter.py:112-121): mask padded frames, right-pad k-1, im2col for Conv1d(C,2C,k,stride 2). */
int vita_whale_adapter_im2col(const void* x, const int32_t* lengths, void* out, int64_t B, int64_t T, int64_t C,
                              int64_t ksize, void* stream);

/* ---- greedy decode step: small kernels --------------------------------------------------------------------- */
/* consume the previous arg-max (best[b]), append it to token_log, advance cache_len, gather its embedding.
 * max_ctx = KV capacity per sequence (pages * page_size): cur_pos saturates at max_ctx - 1.
 * chain_serial (may be NULL): step serial of the completion-counter chain, bumped once per call. */
int vita_decode_embed(uint64_t* best, int32_t* token_log, int32_t* gen_count, int64_t max_log, int32_t* cache_len,
                      int32_t* cur_pos, const void* embed, void* h, int64_t B, int64_t H, int64_t vocab,
                      int64_t max_ctx, uint64_t* chain_serial, void* stream);
/* post_attention_layernorm + router (top-2 of 8) as a stand-alone kernel (the decode step uses the fused form below). */
int vita_decode_router(const void* h, const void* norm_w, const void* gate_w, void* xn, int32_t* topk_ids,
                       float* topk_w, int64_t B, int64_t H, int64_t E, float eps, void* stream);
/* batched decode through the GEMM path: per-sequence KV slots of the current positions, row-wise arg-max of bf16
 * logits into the packed `best` words the decode chain consumes. */
int vita_decode_slots(const int32_t* cur_pos, const int32_t* block_table, int32_t* slots, int64_t B, int64_t page_size,
                      int64_t max_pages, void* stream);
int vita_argmax_rows(const void* logits, uint64_t* best, int64_t B, int64_t V, void* stream);
/* ---- greedy decode step on the tensor cores (tcgen05 swap-AB GEMV, stream-K) ---------------------------------
 * Every linear of the decode step as a weight-streaming GEMV on the tensor cores: the weight tile is the M operand of
 * tcgen05.mma and the activation vector as row 0 of a 16-wide N operand; (row-block, k-block) units are split evenly
 * over all SMs and combined through `workspace` (vita_decode_tc_workspace_bytes, zero-initialised once, shared by all
 * five calls; ws_row_blocks = the max_row_blocks it was sized for): K-partials travel as 64-bit {value, tag} words,
 * the CTA that ends on a row block adds them in slot order (deterministic) and clears the tags.  K multiples of 64.
 * With programmatic dependent launch ("pdl") each kernel triggers its successor once its last weight tile is issued
 * and the successor fills its shared-memory ring before the dependency wait. */
int64_t vita_decode_tc_workspace_bytes(int64_t B, int64_t max_row_blocks);
int vita_decode_tc_qkv_rope(const void* h, const void* norm_w, const void* w_qkv, const float* cos_sin,
                            const int32_t* cur_pos, const int32_t* block_table, void* q_out, void* k_cache,
                            void* v_cache, void* workspace, int64_t ws_row_blocks, int64_t B, int64_t H,
                            int64_t n_q_heads, int64_t n_kv_heads, int64_t head_dim, int64_t page_size,
                            int64_t max_pages, float eps, void* stream);
int vita_decode_tc_oproj(const void* x, const void* w, void* h, void* workspace, int64_t ws_row_blocks, int64_t B,
                         int64_t N, int64_t K, void* stream);
/* route_word ([B] 64-bit words, or NULL) + route_tag (1 .. 2^32-1, different for neighbouring launches that share the
 * word, e.g. layer index + 1): the gate|up kernel publishes {tag, e0, e1} as one word as soon as its router is done and
 * triggers its dependent launch right then; the down projection given the same word and tag streams the rows of the
 * selected experts into its shared-memory ring while the gate|up kernel is still running (it still waits for the
 * activations).  NULL: the down projection reads topk_ids after its dependency wait. */
int vita_decode_tc_moe_gate_up(const void* h, const void* norm_w, const void* gate_w, const void* w13,
                               int32_t* topk_ids, float* topk_w, void* act, void* workspace, int64_t ws_row_blocks,
                               int64_t B, int64_t H, int64_t I, int64_t E, float eps, uint64_t* route_word,
                               int64_t route_tag, void* stream);
int vita_decode_tc_moe_down(const void* act, const void* w2, const int32_t* topk_ids, const float* topk_w, void* h,
                            void* workspace, int64_t ws_row_blocks, int64_t B, int64_t H, int64_t I, int64_t E,
                            const uint64_t* route_word, int64_t route_tag, void* stream);
int vita_tc_lm_head_argmax(const void* h, int64_t h_stride, const void* norm_w, const void* w, void* logits,
                           uint64_t* best, void* workspace, int64_t ws_row_blocks, int64_t B, int64_t H, int64_t V,
                           float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITA_B200_H */
