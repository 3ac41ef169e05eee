/*
 * visper_b200.h — C ABI of libvisper_b200.so, the sm_100a kernel library behind the VisPer-LM
 * (f.k.a. OLA-VLM) data-parallel training step.
 *
 * The reference (SHI-Labs/VisPer-LM @ f7baf3bb) is 100 % Python: it has no FFI of its own, every
 * FLOP is a call into torch / cuBLAS / flash_attn.  This header therefore declares the native
 * layer that sits *beneath* the reference's Python operator surface; each entry point names the
 * reference call site (file:line under /root/reference) whose library call it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are DEVICE pointers owned by the caller
 *     (workspaces included); kernels never allocate, never synchronise, never throw;
 *   - bf16 storage (raw uint16 bit patterns behind `void*`), fp32 accumulation / statistics;
 *   - `ld*` are row strides in ELEMENTS; matrices are row-major; 16-byte aligned, ld % 8 == 0;
 *   - `stream` is a cudaStream_t passed as void*; calls are re-entrant per stream;
 *   - return 0 on success, negative on error; vpb_last_error() gives a thread-local message.
 */
#ifndef VISPER_B200_H
#define VISPER_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VPB_ABI_VERSION 1

/* GEMM epilogue activations */
#define VPB_ACT_NONE 0
#define VPB_ACT_GELU 1       /* erf GELU — nn.GELU() in mm_projector / FeedForward / build_mlp */
#define VPB_ACT_QUICK_GELU 2 /* x*sigmoid(1.702x) — CLIP MLP */
#define VPB_ACT_RELU 3

/* ---- library ------------------------------------------------------------------------------ */
int vpb_abi_version(void);
const char* vpb_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t vpb_launch_count(void);
void vpb_reset_launch_count(void);
/* process-wide kernel-selection switches (testing / A-B timing): key VPB_OPT_*, value 0/1.
   Keys 10-14 default to 1 (measured faster on B200, round 2: profiles/r02_variants_ab.txt); all others to 0. */
#define VPB_OPT_ATTN_LEGACY_FWD 0 /* 1: force the mma.sync attention forward */
#define VPB_OPT_ATTN_LEGACY_BWD 1 /* 1: force the mma.sync attention backward */
#define VPB_OPT_ATTN_TC_BWD_V1 2   /* 1: tcgen05 attention backward without the ping-pong groups */
#define VPB_OPT_GEMM_1CTA 3        /* 1: never use the CTA-pair (cta_group::2) GEMM kernel */
#define VPB_OPT_GEMM_L2_HINTS 8    /* 1: CTA-pair GEMM loads carry L2 eviction hints (A panel evict_last, B evict_first) */
#define VPB_OPT_GEMM_PANEL_MB 4    /* >0: MB of the A operand kept L2-resident per tile-order group (default 32) */
#define VPB_OPT_ATTN_BWD_SS 5      /* 1: attention backward stages P/dS through shared memory (not TMEM) */
#define VPB_OPT_ATTN_BWD_PINGPONG 7 /* 1: attention backward with two softmax groups on alternate iterations (default: column split) */
#define VPB_OPT_ATTN_BWD_DQ_R1 6  /* 1: dQ backward on the round-1 kernel (one CTA per query-tile pair) instead of the persistent kernel */
#define VPB_OPT_ATTN_FWD_NS2 9     /* 1: tcgen05 attention forward with one work item per CTA (round 1) instead of the persistent kernel; 2: the persistent kernel also for launches with few work items (tests) */
#define VPB_OPT_NORM_R1 15          /* 1: RMSNorm fwd/bwd on the round-1 kernels (row as fp32 in registers, 4 / 3 CTAs per SM) instead of the packed high-occupancy ones */
#define VPB_OPT_ATTN_FWD_TC64 12   /* default 1 — non-causal head_dim-64 attention forward (CLIP ViT-L, DINOv2-L towers) on the tcgen05 kernel (one 64-column chunk per tile) instead of the mma.sync kernel */
#define VPB_OPT_GEMM_EPI8 13       /* default 1 — CTA-pair GEMM with EIGHT epilogue warps per CTA (two per TMEM lane quarter, half the columns each) for K <= 1024, where the bias/GELU/store epilogue outlasts the tile's MMAs */
#define VPB_OPT_GATHER_FLAT 14     /* default 1 — vpb_gather_rows for rows of <= 2048 elements as a flat grid-stride loop instead of one CTA per row */
#define VPB_OPT_DWCONV_FFMA2 11    /* default 1 — depthwise 7x7 with packed fp32 FMAs (fma.rn.f32x2 = SASS FFMA2, one per channel pair); bit-identical results, half the FMA instructions */
#define VPB_OPT_WIN_ATTN_V2 10     /* default 1 — vpb_attn_fwd_bias on the one-pass kernel for windows of <= 144 tokens: 1 = two CTAs per SM (96 registers, small spills), 2 = one CTA per SM (no spills) */
int vpb_set_option(int key, int value);
/* profiling aid: device buffer of 16*512 int64 that CTA 0 of the attention dK/dV kernel fills with
 * clock64 stamps of its pipeline events (NULL = off, the default) */
void vpb_set_trace_buffer(void* device_ptr);

/* ---- GEMM: tcgen05 + TMEM + TMA ----------------------------------------------------------
 * C[M,N] = act(A·Bᵀ + bias) + residual ; optional aux = A·Bᵀ + bias (pre-activation copy).
 * a_layout 0: A is [M,K] (K contiguous); 1: A is [K,M] (M contiguous).
 * b_layout 0: B is [N,K] (nn.Linear weight); 1: B is [K,N] (N contiguous).
 * Replaces every torch.nn.functional.linear / cuBLAS GEMM on the path: HF Llama/Phi-3 q/k/v/o,
 * gate/up/down, lm_head (ola_llama.py:105,121), CLIP q/k/v/out/fc1/fc2 + patch conv
 * (clip_encoder.py:56), mm_projector (multimodal_projector/builder.py:56-60), resampler
 * proj_in/to_q/to_kv/to_out/FF/proj_out (resampler.py:9-16,42-44,181,185,213-222), depth-head
 * linear_1..3 (aux_heads/da_v2_head.py:450-455) — forward, dgrad (b_layout 1) and wgrad
 * (a_layout 1, b_layout 1). */
int vpb_gemm_bf16(const void* A, int64_t lda, int a_layout, const void* B, int64_t ldb,
                  int b_layout, void* C, int64_t ldc, int M, int N, int K, int act,
                  const void* bias, const void* residual, int64_t ldr, void* aux, int64_t ldaux,
                  void* stream);

/* Fused SwiGLU epilogues of the MLP GEMMs (HF LlamaMLP / Phi3MLP: down(silu(gate(x))·up(x)),
 * called through ola_llama.py:105).  Wgu is the fused [2F, K] gate|up weight (gate rows first).
 * fwd: gu[M,2F] = A·Wguᵀ (stored when gu != NULL, needed by the backward) and h[M,F] = silu(g)·u in
 *      one launch; F % 128 == 0.
 * bwd: dgu[M,2F] = swiglu'(gu) ∘ (dY·W), W = down_proj weight: b_layout 1 → [K, F] as stored by
 *      nn.Linear (read MN-major), b_layout 0 → a K-major transposed copy [F, K].
 * gu_tiled 1: g|u is kept in the tile-major layout [M/128][F/32][128 rows][32 gate | 32 up]
 *      (ceil(M/128)*128*2F elements, ldgu ignored) that makes the row-owning epilogue threads touch
 *      whole 128-byte lines; only the two fused kernels read it. */
int vpb_gemm_swiglu_fwd(const void* A, int64_t lda, const void* Wgu, int64_t ldw, void* gu,
                        int64_t ldgu, int gu_tiled, void* h, int64_t ldh, int M, int F, int K,
                        void* stream);
int vpb_gemm_swiglu_bwd(const void* dY, int64_t lddy, const void* W, int64_t ldw, int b_layout,
                        const void* gu, int64_t ldgu, int gu_tiled, void* dgu, int64_t lddgu, int M,
                        int F, int K, void* stream);

/* Fused QKV projection + rotary embedding (HF LlamaAttention q/k/v_proj + apply_rotary_pos_emb):
 * C = A·Bᵀ with the first rope_heads 128-wide heads of every row rotated by the row's position
 * (pos_ids[row], or row % seq_len when NULL) in the GEMM epilogue.  head_dim 128, N % 256 == 0. */
int vpb_gemm_rope_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                       int M, int N, int K, const float* cos_t, const float* sin_t, int seq_len,
                       const int* pos_ids, int rope_heads, void* stream);

/* ---- normalisation -------------------------------------------------------------------------
 * HF LlamaRMSNorm / Phi3RMSNorm (eps 1e-5) and nn.LayerNorm (CLIP, resampler.py). */
int vpb_rmsnorm_fwd(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, float* rstd,
                    int M, int D, float eps, void* stream);
/* dx = rmsnorm'(dy) (+ dres): fuses the residual-stream gradient add */
int vpb_rmsnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const void* w,
                    const float* rstd, const void* dres, int64_t lddres, void* dx, int64_t lddx,
                    int M, int D, void* stream);
int vpb_layernorm_fwd(const void* x, int64_t ldx, const void* w, const void* b, void* y,
                      int64_t ldy, float* mean, float* rstd, int M, int D, float eps, void* stream);
int vpb_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const void* w,
                      const float* mean, const float* rstd, const void* dres, int64_t lddres,
                      void* dx, int64_t lddx, int M, int D, void* stream);
/* out[n] = sum_m a[m,n] * (b ? (b[m,n]-mean[m])*rstd[m] : 1) — bias / norm-weight gradients */
int vpb_colsum(const void* a, int64_t lda, const void* b, int64_t ldb, const float* mean,
               const float* rstd, float* out, int M, int N, void* stream);

/* ---- elementwise ----------------------------------------------------------------------------
 * RoPE: HF apply_rotary_pos_emb (rotate_half convention), applied in place to `nheads` heads of
 * the packed QKV buffer; inverse=1 is the backward. SwiGLU: LlamaMLP / Phi3MLP silu(gate)*up on
 * a packed [M, 2F] gate|up buffer. */
int vpb_rope_table(float* cos_t, float* sin_t, int max_pos, int head_dim, float theta, void* stream);
int vpb_rope_inplace(void* x, int64_t ld, int M, int seq_len, const int* pos_ids,
                     const float* cos_t, const float* sin_t, int nheads, int head_dim, int inverse,
                     void* stream);
int vpb_swiglu_fwd(const void* gu, int64_t ldgu, void* h, int64_t ldh, int M, int F, void* stream);
int vpb_swiglu_bwd(const void* gu, int64_t ldgu, const void* dh, int64_t lddh, void* dgu,
                   int64_t lddgu, int M, int F, void* stream);
int vpb_act_bwd(const void* pre, int64_t ldp, const void* dy, int64_t lddy, void* dx, int64_t lddx,
                int M, int N, int act, void* stream);
int vpb_axpby(const void* a, const void* b, void* out, float alpha, float beta, int64_t n,
              void* stream);
/* out = in * (*scale), scalar read on the device */
int vpb_scale_dev(const void* in, void* out, const float* scale, int64_t n, void* stream);
int vpb_transpose(const void* in, int64_t ldi, void* out, int64_t ldo, int R, int C, void* stream);
int vpb_cast_f32_bf16(const float* in, void* out, int64_t n, float scale, void* stream);

/* ---- CLIP patch embedding (clip_encoder.py:56 → HF CLIPVisionEmbeddings) -------------------- */
int vpb_im2col_patches(const void* images, void* out, int B, int H, int W, int patch, int Kpad,
                       void* stream);
int vpb_clip_embed(const void* patch, const void* cls, const void* pos, void* out, int B,
                   int npatch, int D, void* stream);

/* ---- frozen DPT depth decoder (aux_heads/da_v2_head.py:181-321 → `depth_preds`) --------------
 * NHWC bf16 activations; every conv is im2col + vpb_gemm_bf16 (K order ky,kx,c). */
int vpb_im2col3x3_nhwc(const void* in, void* out, int B, int H, int W, int C, int stride, int relu_in,
                       void* stream);
/* F.interpolate(mode="bilinear", align_corners=True) */
int vpb_bilinear_nhwc(const void* in, void* out, int B, int Hi, int Wi, int Ho, int Wo, int C,
                      void* stream);
/* F.interpolate(mode="bilinear", align_corners=False) — the 25x25 → 24x24 resize of the seg teacher's
 * last feature map (aux_heads/oneformer_head.py:31) */
int vpb_bilinear_nhwc_half_pixel(const void* in, void* out, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                 void* stream);
/* ConvTranspose2d(kernel = stride = k) epilogue: [B*H*W, k*k*C] → [B, H*k, W*k, C] (+bias) */
int vpb_pixel_shuffle_nhwc(const void* in, const void* bias, void* out, int B, int H, int W, int C,
                           int k, void* stream);
/* 1x1 conv to a single channel (+ReLU), fp32 output */
int vpb_conv1x1_to1(const void* in, const void* w, const void* bias, float* out, int64_t P, int C,
                    int relu, void* stream);
/* per-image (x - min) / (max - min), base_ola_vlm.py:466-469 */
int vpb_minmax_normalize(const float* in, float* out, int B, int64_t n, void* stream);

/* ---- ConvNeXt tower (multimodal_encoder/clip_convnext_encoder.py:150-174 → timm ConvNeXt block) ----
 * depthwise Conv2d(C, C, 7, padding=3, groups=C) on NHWC bf16: out[b,y,x,c] = bias[c] +
 * sum_{ky,kx} in[b, y+ky-3, x+kx-3, c] * w49[ky*7+kx, c]  (zero outside), fp32 accumulate.
 * w49 is the [C,1,7,7] filter repacked tap-major [49, C]; C % 64 == 0; out must not alias in. */
int vpb_dwconv7x7_nhwc(const void* in, const void* w49, const void* bias, void* out, int B, int H, int W,
                       int C, void* stream);

/* ---- multimodal splice (ola_arch.py:256-444 prepare_inputs_labels_for_multimodal) -----------
 * One gather from a host-built index plan replaces the per-sample Python cat loop. */
int vpb_gather_rows(void* out, int64_t ldo, int nrows, int D, const int* kind, const int* index,
                    const void* src0, int64_t ld0, const void* src1, int64_t ld1, const void* src2,
                    int64_t ld2, const void* src3, int64_t ld3, void* stream);
int vpb_gather_sum_rows(void* out, int64_t ldo, int nslots, int cnt, const int* index,
                        const void* src, int64_t lds, int D, float scale, void* stream);
int vpb_scatter_add_rows(float* dst, int64_t ldd, int nrows, const int* index, const void* src,
                         int64_t lds, int D, void* stream);
/* dst[index[r]] += src[r] (bf16; indices unique per call) */
int vpb_add_rows(void* dst, int64_t ldd, int nrows, const int* index, const void* src, int64_t lds,
                 int D, void* stream);
/* task-token pooling param[576,D].view(8,72,D).mean(1) (ola_arch.py:225-228) */
int vpb_group_mean(const void* in, int64_t ldi, void* out, int64_t ldo, int groups, int gsize,
                   int D, void* stream);
int vpb_group_mean_bwd(const void* dout, int64_t ldo, void* din, int64_t ldi, int groups,
                       int gsize, int D, void* stream);

/* ---- attention ------------------------------------------------------------------------------
 * Flash attention on packed projections; optional second K/V segment (k2/v2, length sk2) is
 * appended after the first (PerceiverAttention keys = cat(x, latents), resampler.py:61-62).
 * lse: [B,H,sq] fp32. delta: [B,H,sq] fp32 workspace for the backward.
 * window > 0 (causal only): sliding-window attention as HF 4.41.1 Phi3FlashAttention2 passes it to
 * flash_attn (window_size=(sliding_window, sliding_window)): key j is visible to query i iff
 * 0 <= i - j <= window.  Phi-3-mini: sliding_window 2047. */
int vpb_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                 const void* k2, int64_t ldk2, const void* v2, int64_t ldv2, void* o, int64_t ldo,
                 float* lse, int B, int H, int KVH, int sq, int sk, int sk2, int head_dim,
                 float scale, int causal, int window, void* stream);
int vpb_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                 const void* k2, int64_t ldk2, const void* v2, int64_t ldv2, const void* o,
                 int64_t ldo, const void* dO, int64_t lddo, const float* lse, float* delta, void* dq,
                 int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, void* dk2,
                 int64_t lddk2, void* dv2, int64_t lddv2, int B, int H, int KVH, int sq, int sk,
                 int sk2, int head_dim, float scale, int causal, int window, void* stream);
/* Forward-only window attention with an additive score bias (frozen Swin teacher: HF
 * SwinSelfAttention.forward — scores = q.k/sqrt(hd) + relative_position_bias[h] + attn_mask[window]):
 * bias fp32 [H, sq, sk]; bias_mask fp32 [mask_mod, sq, sk] or NULL, batch entry b uses mask b % mask_mod
 * (windows are batch-major: b = image * n_windows + window).  head_dim 32, non-causal, KVH == H. */
int vpb_attn_fwd_bias(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                      void* o, int64_t ldo, float* lse, int B, int H, int sq, int sk, int head_dim,
                      float scale, const float* bias, const float* bias_mask, int mask_mod, void* stream);
/* Self-attention backward (sq == sk == seq_len, one K/V segment) that also undoes the rotary
 * embedding: dq / dk come back as gradients w.r.t. the PRE-rotation projections, i.e. what
 * vpb_attn_bwd followed by vpb_rope_inplace(inverse=1) on dq (H heads) and dk (KVH heads) returns,
 * bit for bit (the backward of HF apply_rotary_pos_emb, modeling_llama.py).  head_dim 128 runs the
 * rotation in the tcgen05 backward epilogues; other shapes fall back to the separate rope kernel.
 * cos_t / sin_t: vpb_rope_table output; pos_ids: int32 [B*seq_len] or NULL (position = t). */
int vpb_attn_bwd_rope(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                      int64_t ldv, const void* o, int64_t ldo, const void* dO, int64_t lddo,
                      const float* lse, float* delta, void* dq, int64_t lddq, void* dk, int64_t lddk,
                      void* dv, int64_t lddv, int B, int H, int KVH, int seq_len, int head_dim,
                      float scale, int causal, int window, const float* cos_t, const float* sin_t,
                      const int* pos_ids, void* stream);

/* ---- next-token cross-entropy (ola_llama.py:121-136) -----------------------------------------
 * labels: int64 [B,T] UNSHIFTED when shift=1 (row (b,t) is scored against labels[b,t+1]).
 * vpb_ce_fwd_bwd handles rows row0..row0+R-1 of the flattened [B*T] sequence whose logits sit in
 * `logits` (bf16, overwritten with gscale*(softmax-onehot)/count when write_grad). */
int vpb_ce_count(const int64_t* labels, int64_t R, int T, int shift, float* count_out, void* stream);
int vpb_ce_fwd_bwd(void* logits, int64_t ld, const int64_t* labels, int64_t row0, int R, int V,
                   int T, int shift, float* row_loss, const float* count, float gscale,
                   int write_grad, void* stream);
int vpb_ce_finalize(const float* row_loss, int64_t R, const float* count, float* loss_out,
                    void* stream);

/* ---- embedding-distillation loss (base_ola_vlm.py:289-320, ola_utils.py:108-125) --------------
 * pred [B,n], tgt [Bt,n] (all-gathered targets, own rows start at `off`), tau = logit_scale
 * parameter, mask [B] fp32 or NULL. out4 = {loss, smooth_l1, contrastive, dloss/dtau}.
 * coef: 2B + B*Bt floats consumed by the backward; stats (optional): B*Bt dots | B |p|² | Bt |t|² | B sl1. */
int64_t vpb_distill_workspace_floats(int B, int Bt, int64_t n);
int vpb_distill_loss_fwd(const void* pred, int64_t ldp, const void* tgt, int64_t ldt, int64_t n,
                         int B, int Bt, int off, const float* tau, const float* mask,
                         float contrastive_weight, float* workspace, float* out4, float* coef,
                         float* stats, void* stream);
int vpb_distill_loss_bwd(const void* pred, int64_t ldp, const void* tgt, int64_t ldt, int64_t n,
                         int B, int Bt, int off, const float* coef, const float* gout, void* dpred,
                         int64_t lddp, void* stream);

/* ---- sharded optimizer (HF adamw_torch under DeepSpeed ZeRO-2, scripts/zero2.json) ------------ */
int vpb_adamw_step(float* master, float* m, float* v, const void* grad, void* param, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   const float* grad_scale, void* stream);
int vpb_grad_sumsq(const void* grad, int64_t n, float* workspace, float* out, int accumulate,
                   void* stream);
int vpb_clip_coef(const float* sumsq, float max_norm, float extra_scale, float* coef,
                  float* norm_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VISPER_B200_H */
