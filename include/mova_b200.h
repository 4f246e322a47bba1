/*
 * mova_b200 -- C ABI of the sm_100a kernels behind the MOVA dual-tower DiT denoising-block forward.
 *
 * The reference (Jp-17/DualForce == OpenMOSS/MOVA) has no FFI: its hot path is Python nn.Modules whose
 * GPU work is done by library kernels.  Each entry point below replaces the library call(s) made at the
 * cited reference lines; the Python host (dualforce_b200/) binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch); nothing is allocated or retained,
 *   - activations and weights are bf16 (reference model dtype, scripts/inference_single.py:77),
 *     modulation / RoPE tables / LSE are fp32,
 *   - strides (ld*) are in ELEMENTS; rows must be 16-byte aligned (ld % 8 == 0 for bf16),
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no host synchronisation,
 *   - return 0 on success, negative on error; mova_b200_last_error() describes the last failure of
 *     the calling thread.  There is no CPU fallback: a non-sm_100 device is an error.
 */
#ifndef MOVA_B200_H_
#define MOVA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOVA_B200_ABI_VERSION 6

/* epilogues of mova_b200_linear */
#define MOVA_EPI_BIAS 0      /* C = A W^T + b                          nn.Linear                          */
#define MOVA_EPI_GELU_TANH 1 /* C = gelu_tanh(A W^T + b)               wan_video_dit.py:270-271 (ffn.0+1) */
#define MOVA_EPI_RESIDUAL 2  /* C = R + gate[n] * scale * (A W^T + b)  wan_video_dit.py:254-255,287-290;  */
                             /*                                        interactionv2.py:535               */

/* RoPE conventions of mova_b200_rmsnorm_rope */
#define MOVA_ROPE_NONE 0
#define MOVA_ROPE_INTERLEAVED 1 /* (2i,2i+1) complex pairs, table [L, head_dim/2]  wan_video_dit.py:131-137 */
#define MOVA_ROPE_HALF 2        /* rotate-half (i, i+hd/2), table [L, head_dim]    interactionv2.py:40-72   */

int mova_b200_abi_version(void);
const char* mova_b200_last_error(void);

/*
 * Diagnostics: 16 host-visible words a kernel fills in before trapping on a pipeline time-out
 * {0x4d564442, mbarrier smem address, parity, blockIdx.x/y/z, threadIdx.x, 0...}; word 0 is 0 when nothing tripped.
 */
const uint32_t* mova_b200_debug_record(void);

/* 0 if `device` is an sm_100 part the kernels can run on, negative otherwise. */
int mova_b200_device_check(int device);

/*
 * nn.Linear with fused epilogue: C[M,N] = epi(A[M,K] . W[N,K]^T + bias[N]).
 * Replaces cuBLAS GEMM + ATen elementwise at wan_video_dit.py:171-174,218-221,270-271,287-290 and
 * interactionv2.py:218-221,251,535.  tcgen05 (UMMA 128x256x16 or CTA-pair 256x256x16), TMA-fed,
 * persistent, fp32 accumulation in TMEM.
 *   bias      bf16 [N] or NULL
 *   residual  bf16 [M, ldr] (MOVA_EPI_RESIDUAL only; may alias C)
 *   gate      fp32 [N] or NULL (=1)           scale: scalar multiplier of the gated branch
 *   cta_group 1 or 2 (2 = CTA-pair UMMA, M tile 256); 0 = library default
 */
int mova_b200_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                     int64_t ldc, int M, int N, int K, int epilogue, const void* residual, int64_t ldr,
                     const float* gate, float scale, int cta_group, void* stream);

/*
 * Same, with segmented operands for the context-parallel (Ulysses) layouts, so no pack / unpack copy is needed
 * around the all-to-all (the reference's yunchang LongContextAttention, called from wan_video_dit.py:207, permutes
 * and copies on both sides):
 *   A(m, k) = A[(k / seg_k) * a_seg_stride + m * lda + (k % seg_k)]      seg_k % 64 == 0 or seg_k == K
 *       -- what the inverse all-to-all delivers: [source rank][token][heads of that rank];
 *   C(m, n) = C[(n / seg_n) * c_seg_stride + m * ldc + (n % seg_n)]      seg_n % 64 == 0 or seg_n == N
 *       -- what the forward all-to-all sends: [destination rank][token][q|k|v heads of that rank]
 *          (MOVA_EPI_RESIDUAL is not available with a segmented C).
 */
int mova_b200_linear_ex(const void* A, int64_t lda, int seg_k, int64_t a_seg_stride, const void* W, int64_t ldw,
                        const void* bias, void* C, int64_t ldc, int seg_n, int64_t c_seg_stride, int M, int N, int K,
                        int epilogue, const void* residual, int64_t ldr, const float* gate, float scale,
                        int cta_group, void* stream);

/*
 * Non-causal softmax attention, head_dim 128: O = softmax(Q K^T * softmax_scale) V.
 * Replaces flash_attention() at wan_video_dit.py:58-91 (called from :188, :241, interactionv2.py:250).
 *   q: [B, Sq, H, 128]  element (b,s,h,d) at q + b*q_bs + s*q_ss + h*128 + d   (same for k, v, o)
 *   lse: fp32 [B, H, Sq] natural-log-sum-exp of the scaled scores, or NULL
 */
int mova_b200_attn_fwd(const void* q, int64_t q_bs, int64_t q_ss, const void* k, int64_t k_bs, int64_t k_ss,
                       const void* v, int64_t v_bs, int64_t v_ss, void* o, int64_t o_bs, int64_t o_ss,
                       float* lse, int B, int Sq, int Skv, int H, int D, float softmax_scale, void* stream);

/*
 * Development / measurement entry: the same attention with the schedule chosen explicitly.
 *   variant 92: round-2 schedule, CTA pair (cta_group::2)  -- what mova_b200_attn_fwd runs for long key sequences
 *   variant  3: round-1 schedule (two query tiles per CTA)   -- what it runs for <= 8 key blocks against many queries
 *   emu: share of the exponentials evaluated as a polynomial on the FMA pipe, in 16ths of the score pairs (4)
 *   trace: NULL, or device memory for 3 x 4096 u64 event records (clock << 8 | id) of CTA (0,0,0)
 */
int mova_b200_attn_fwd_variant(const void* q, int64_t q_bs, int64_t q_ss, const void* k, int64_t k_bs, int64_t k_ss,
                               const void* v, int64_t v_bs, int64_t v_ss, void* o, int64_t o_bs, int64_t o_ss,
                               float* lse, int B, int Sq, int Skv, int H, int D, float softmax_scale, int variant,
                               int emu, unsigned long long* trace, void* stream);

/*
 * Merge `n_parts` partial attention results over disjoint key sets (split-KV / context-parallel v2a):
 *   o = sum_p exp(lse_p - lse) o_p,  lse = log sum_p exp(lse_p).
 *   o_parts: bf16 [n_parts, rows, H*D] contiguous, lse_parts: fp32 [n_parts, H, rows]
 *   out: bf16 [rows, ldo], lse_out fp32 [H, rows] or NULL
 */
int mova_b200_lse_merge(const void* o_parts, const float* lse_parts, int n_parts, void* out, int64_t ldo,
                        float* lse_out, int rows, int H, int D, void* stream);

/*
 * y = LayerNorm(x) [* ln_w + ln_b] [* (1 + scale) + shift], one pass, fp32 statistics.
 * Replaces nn.LayerNorm + modulate() at wan_video_dit.py:94-96,267-269,286,289 and y_norm at
 * interactionv2.py:322,349.   ln_w/ln_b: bf16 [d] or NULL;  shift/scale: fp32 [d] or NULL.
 */
int mova_b200_layernorm(const void* x, int64_t ldx, void* y, int64_t ldy, int L, int d, float eps,
                        const void* ln_w, const void* ln_b, const float* shift, const float* scale,
                        void* stream);

/*
 * In-place x = RoPE(RMSNorm_full(x) * w): RMS over the whole row of d = H*head_dim channels
 * (torch.nn.RMSNorm(dim), wan_video_dit.py:175-176,181-182; interactionv2.py:222-223,229-230) followed by
 * the rotary embedding of wan_video_dit.py:131-137 (interleaved) or interactionv2.py:47-72 (rotate-half).
 *   cos/sin: fp32 tables, row l for token l (see MOVA_ROPE_*), NULL when rope_mode == MOVA_ROPE_NONE
 */
int mova_b200_rmsnorm_rope(void* x, int64_t ldx, int L, int d, int head_dim, const void* w, float eps,
                           const float* cos_tab, const float* sin_tab, int rope_mode, void* stream);

/*
 * Same, for a row stored as d/seg_len segments of seg_len channels (a multiple of 128) that live `seg_stride`
 * elements apart: channel c of token l is x[(c / seg_len) * seg_stride + l * ldx + (c % seg_len)].  Used on the
 * destination-rank-major q/k/v buffer of the context-parallel path, before its all-to-all.
 */
int mova_b200_rmsnorm_rope_seg(void* x, int64_t ldx, int seg_len, int64_t seg_stride, int L, int d, int head_dim,
                               const void* w, float eps, const float* cos_tab, const float* sin_tab, int rope_mode,
                               void* stream);

/* out[i] = float(a[i]) + float(b[i]), bf16 inputs (b may be NULL): modulation + t_mod, wan_video_dit.py:279-280 */
int mova_b200_add_to_f32(const void* a, const void* b, float* out, int64_t n, void* stream);

/* ---- the step either side of the dual-tower forward: MOVA.inference_single_step, pipeline_mova.py:500-609 ---- */

/*
 * im2col of the patch embedding (nn.Conv3d / nn.Conv1d with stride == kernel: wan_video_dit.py:367-368,399-409;
 * wan_audio_dit.py:143-145,180-189) so the convolution runs as one mova_b200_linear call on
 * weight.view(dim, C*pt*ph*pw).  Also replaces the `.to(model_dtype)` cast at pipeline_mova.py:556-557.
 *   x:   [C, F, H, W] contiguous, fp32 (x_is_f32 != 0) or bf16     (audio: H = W = 1, patch (p, 1, 1))
 *   out: bf16 [L, ldo], L = (F/pt)(H/ph)(W/pw) tokens in (f, h, w) order, column ((c*pt + dt)*ph + dh)*pw + dw
 */
int mova_b200_patchify(const void* x, int x_is_f32, int C, int F, int H, int W, int pt, int ph, int pw, void* out,
                       int64_t ldo, void* stream);

/*
 * WanModel.unpatchify / WanAudioModel.unpatchify (wan_video_dit.py:411-416; wan_audio_dit.py:191-195):
 *   in:  bf16 [L, ldi] head output, column ((x*ph + y)*pw + z)*Cout + c, tokens in (f, h, w) order
 *   out: bf16 [Cout, Fp*pt, Hp*ph, Wp*pw] contiguous
 */
int mova_b200_unpatchify(const void* in, int64_t ldi, void* out, int Cout, int Fp, int Hp, int Wp, int pt, int ph,
                         int pw, void* stream);

/*
 * sinusoidal_embedding_1d (wan_video_dit.py:99-103), fp64 math like the reference:
 *   out[i] = cos(t[0] * 10000^(-i/(dim/2))), out[dim/2 + i] = sin(...), i < dim/2;  t: one fp32 value on the device
 */
int mova_b200_sinusoidal(const float* t, float* out, int dim, void* stream);

/*
 * One row of an nn.Linear in fp32: y[n] = post(sum_k pre(x[k]) W[n,k] + bias[n]); pre/post: 0 = identity, 1 = SiLU.
 * The time_embedding / time_projection MLPs (wan_video_dit.py:374-380) which the reference evaluates under
 * autocast(float32) (pipeline_mova.py:544-549): fp32 activations, bf16-valued weights.
 *   x fp32 [K], W bf16 [N, ldw], bias bf16 [N] or NULL, y fp32 [N], y_bf16 bf16 [N] or NULL (rounded copy)
 */
int mova_b200_gemv_f32(const float* x, const void* W, int64_t ldw, const void* bias, float* y, void* y_bf16, int N,
                       int K, int pre_act, int post_act, void* stream);

/*
 * Classifier-free guidance + flow-match Euler update in one pass:
 *   out = sample + (nega + cfg_scale * (posi - nega)) * dsigma        (nega == NULL: out = sample + posi * dsigma)
 * Replaces `nega.float() + cfg_scale * (posi.float() - nega.float())` (pipeline_mova.py:456-460) and
 * FlowMatchPairScheduler.step_from_to's `sample + model_output * (sigma_to - sigma_from)`
 * (schedulers/flow_match_pair.py:213-227); dsigma = sigma_to - sigma_from comes from the reference's scheduler.
 *   posi/nega bf16 [n], sample/out fp32 [n] (out may alias sample); all 16-byte aligned
 */
int mova_b200_cfg_euler(const void* posi, const void* nega, const float* sample, float* out, int64_t n,
                        float cfg_scale, float dsigma, void* stream);

/* ---- context-parallel exchange over NVSwitch peer memory (no NCCL kernel on the data path) ---- */

/*
 * The Ulysses head <-> sequence all-to-all that yunchang's LongContextAttention performs inside USPAttention.forward
 * (wan_video_dit.py:192-208; `dist.all_to_all_single` underneath) as copy-engine transfers into the peers' receive
 * windows plus flag words, so the exchange of one head group proceeds while the attention kernel of the previous
 * group occupies every SM (an NCCL all-to-all is a kernel and waits for SMs: measured 0.75-2.1 ms instead of
 * 0.10-0.16 ms per head group at cp = 8).  These are the only entry points that own memory.
 *
 *   peer_alloc   cudaMalloc + zero-fill a window on the current device and export its 64-byte cudaIpcMemHandle_t
 *   peer_open    map a window exported by another process of this node (enables peer access lazily);
 *                the returned pointer is valid on the current device
 *   peer_close / peer_free   undo peer_open / peer_alloc
 */
int mova_b200_peer_alloc(int64_t nbytes, void** ptr, void* handle64);
int mova_b200_peer_open(const void* handle64, void** ptr);
int mova_b200_peer_close(void* ptr);
int mova_b200_peer_free(void* ptr);

/* 1 when the driver offers 64-bit stream memory operations (cuStreamWriteValue64 / cuStreamBatchMemOp) on the current device */
int mova_b200_peer_memops_supported(void);

/*
 * Enqueue on `stream`: n_copies x cudaMemcpyAsync(dst[i], src[i], nbytes[i]) (local or peer-mapped device pointers,
 * contiguous chunks), then store `epoch` (> 0, increasing over the life of the windows) into each of the n_flags
 * (<= 32) 8-byte flag words -- typically one word in every destination's window.  A consumer that observes the flag
 * observes the copied bytes.
 *   epoch_src != NULL  no kernel: cuStreamWriteValue64 puts the epoch into *epoch_src (a device word of the caller; all
 *                      pushes sharing it must be queued on ONE stream), then one 8-byte peer copy per flag word carries
 *                      it; flag i with bit i of local_mask set lives in THIS device's memory and is written directly
 *                      with cuStreamWriteValue64 (a same-device copy would run on the SMs)
 *   epoch_src == NULL  one 32-thread kernel stores the flags (st.release.sys) -- needs an SM, so it queues behind a
 *                      kernel that occupies the whole device
 */
int mova_b200_peer_push(int n_copies, void* const* dst, const void* const* src, const int64_t* nbytes, int n_flags,
                        void* const* flags, int64_t local_mask, int64_t epoch, void* epoch_src, void* stream);

/*
 * Enqueue on `stream` a wait until all n_flags consecutive 8-byte words at `flags` (this device's own window) hold a
 * value >= epoch.
 *   use_memops != 0  cuStreamBatchMemOp(WAIT_VALUE_64, GEQ): no kernel, no time-out
 *   use_memops == 0  one polling kernel (ld.acquire.sys); after timeout_ms without progress it writes
 *                    {0x4d565057, flag index, epoch, value seen} to mova_b200_debug_record() and traps: a lost peer
 *                    becomes a CUDA error on this rank instead of a silent wait
 */
int mova_b200_peer_wait(const void* flags, int n_flags, int64_t epoch, int timeout_ms, int use_memops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOVA_B200_H_ */
