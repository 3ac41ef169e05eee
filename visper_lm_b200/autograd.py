"""torch.autograd.Functions that sequence the C-ABI kernels for each fused block of the step.

Every Function works on 2-D row-major [rows, features] bf16 tensors (rows = batch·sequence).
Backward passes are written by hand (no torch compute): dgrad / wgrad are the same tcgen05 GEMM
with MN-major operand descriptors, the residual-stream gradient add is fused into the norm
backward, and activations that are cheap to rebuild (norm outputs, SwiGLU products) are
recomputed instead of saved.

Reference call sites are cited per Function; all paths are under /root/reference/ola_vlm/.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, BF16


# =================================================================================================
class LinearFn(torch.autograd.Function):
    """y = act(x·Wᵀ + b) — nn.Linear (+GELU/ReLU) of mm_projector
    (model/multimodal_projector/builder.py:53-60) and the depth-head MLPs
    (model/aux_heads/da_v2_head.py:331-335, 450-455)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        need_pre = act != ACT_NONE and (x.requires_grad or w.requires_grad)
        if need_pre:
            y, pre = ops.gemm(x, w, bias=b, act=act, want_pre=True)
        else:
            y, pre = ops.gemm(x, w, bias=b, act=act), None
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, w, pre)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, pre = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act != ACT_NONE:
            dy = ops.act_bwd(pre, dy, ctx.act)
        dx = ops.gemm(dy, w, b_layout=1) if ctx.needs_input_grad[0] else None
        dw = ops.gemm(dy, x, a_layout=1, b_layout=1) if ctx.needs_input_grad[1] else None
        db = ops.colsum(dy) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None


def linear(x, w, b=None, act=ACT_NONE):
    return LinearFn.apply(x, w, b, act)


# =================================================================================================
class RMSNormFn(torch.autograd.Function):
    """Final model.norm (HF LlamaRMSNorm) — language_model/ola_llama.py:105 → hidden_states[-1]."""

    @staticmethod
    def forward(ctx, x, w, eps):
        y, rstd = ops.rmsnorm_fwd(x, w, eps)
        ctx.save_for_backward(x, w, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.rmsnorm_bwd(dy, x, w, rstd) if ctx.needs_input_grad[0] else None
        dw = ops.colsum(dy, x, None, rstd) if ctx.needs_input_grad[1] else None
        return dx, dw, None


# =================================================================================================
class DecoderLayerFn(torch.autograd.Function):
    """One HF LlamaDecoderLayer / Phi3DecoderLayer (called from language_model/ola_llama.py:105,
    ola_phi3.py): RMSNorm → fused QKV GEMM → RoPE in place → causal GQA flash attention → o_proj
    (+residual in the GEMM epilogue) → RMSNorm → gate|up GEMM → SwiGLU → down_proj (+residual).

    `wq, wk, wv` / `wg, wu` are the reference-named parameters (row views of the fused buffers in
    `meta`; for Phi-3 wq / wg are the already-fused qkv_proj / gate_up_proj and the others None).
    They are autograd inputs so freezing by name works; the kernels read the fused memory.
    """

    @staticmethod
    def forward(ctx, x, n1, wq, wk, wv, wo, n2, wg, wu, wd, meta):
        B, T, H, KVH, hd, eps = meta.B, meta.T, meta.H, meta.KVH, meta.hd, meta.eps
        wqkv, wgu = meta.wqkv, meta.wgu
        qw, kw = H * hd, KVH * hd
        h1, rstd1 = ops.rmsnorm_fwd(x, n1, eps)
        if ops.FUSE_ROPE and hd == 128 and wqkv.shape[0] % 256 == 0:
            qkv = ops.gemm_rope(h1, wqkv, T, meta.cos, meta.sin, H + KVH, pos_ids=meta.pos_ids)  # RoPE in the epilogue
        else:
            qkv = ops.gemm(h1, wqkv)
            ops.rope_(qkv, T, meta.cos, meta.sin, H + KVH, hd, pos_ids=meta.pos_ids)
        del h1
        o, lse = ops.attn_fwd(qkv[:, :qw], qkv[:, qw:qw + kw], qkv[:, qw + kw:], B, H, KVH, T, T, hd,
                              hd ** -0.5, True, window=getattr(meta, "window", 0))
        x2 = ops.gemm(o, wo, residual=x)
        h2, rstd2 = ops.rmsnorm_fwd(x2, n2, eps)
        # g|u is saved tile-major when only the fused backward will read it (down_proj frozen)
        gu_tiled = (ops.FUSE_SWIGLU and ops.FUSE_SWIGLU_BWD and wgu.shape[0] % 256 == 0
                    and not wd.requires_grad)
        if ops.FUSE_SWIGLU and wgu.shape[0] % 256 == 0:
            hh, gu = ops.gemm_swiglu_fwd(h2, wgu, tiled=gu_tiled)  # SwiGLU in the GEMM epilogue
        else:
            gu = ops.gemm(h2, wgu)
            hh = ops.swiglu_fwd(gu)
        del h2
        x3 = ops.gemm(hh, wd, residual=x2)
        ctx.meta = meta
        ctx.split_qkv = wk is not None
        ctx.split_gu = wu is not None
        ctx.gu_tiled = gu_tiled
        ctx.F = wgu.shape[0] // 2
        ctx.save_for_backward(x, n1, wo, n2, wd, rstd1, qkv, o, lse, x2, rstd2, gu)
        return x3

    @staticmethod
    def backward(ctx, dx3):
        x, n1, wo, n2, wd, rstd1, qkv, o, lse, x2, rstd2, gu = ctx.saved_tensors
        meta = ctx.meta
        B, T, H, KVH, hd, eps = meta.B, meta.T, meta.H, meta.KVH, meta.hd, meta.eps
        wqkv, wgu = meta.wqkv, meta.wgu
        qw, kw = H * hd, KVH * hd
        F = ctx.F
        dx3 = dx3.contiguous()
        nig = ctx.needs_input_grad
        g = [None] * 11

        def dgrad(dy, w, wt):  # dY·W: K-major frozen copy Wᵀ when available, else MN-major descriptor
            return ops.gemm(dy, wt) if wt is not None else ops.gemm(dy, w, b_layout=1)

        sink = getattr(meta, "grad_sink", None)

        def wgrad(key, dy, xin):
            """dYᵀ·X.  With a gradient sink (ZeRO-2 optimizer) the GEMM writes — or accumulates into — the
            optimizer's gradient buffer and autograd gets None; otherwise the gradient tensor is returned."""
            d = sink.dst(key) if sink is not None else None
            if d is None:
                return ops.gemm(dy, xin, a_layout=1, b_layout=1)
            buf, acc = d
            ops.gemm(dy, xin, a_layout=1, b_layout=1, out=buf, residual=buf if acc else None)
            sink.done(key)
            return None

        # ---- MLP: x3 = x2 + down(swiglu(gate_up(rmsnorm(x2)))) ----
        if nig[9]:
            hh = ops.swiglu_fwd(gu)
            g[9] = wgrad("d", dx3, hh)
            del hh
        if ctx.gu_tiled:  # dgrad of down_proj with the SwiGLU derivative in its epilogue
            dgu = (ops.gemm_swiglu_bwd(dx3, meta.wdT, gu, b_layout=0, tiled=True, F=F)
                   if meta.wdT is not None
                   else ops.gemm_swiglu_bwd(dx3, wd, gu, b_layout=1, tiled=True, F=F))
        else:
            dh = dgrad(dx3, wd, meta.wdT)
            dgu = ops.swiglu_bwd(gu, dh)
            del dh
        dn2 = dgrad(dgu, wgu, meta.wguT)
        if nig[7] or nig[8]:
            h2, _ = ops.rmsnorm_fwd(x2, n2, eps)
            dwgu = wgrad("gu", dgu, h2)
            del h2
            if dwgu is None:
                pass
            elif ctx.split_gu:
                g[7], g[8] = dwgu[:F], dwgu[F:]
            else:
                g[7] = dwgu
        del dgu
        if nig[6]:
            g[6] = ops.colsum(dn2, x2, None, rstd2)
        dx2 = ops.rmsnorm_bwd(dn2, x2, n2, rstd2, dres=dx3)  # fused residual-gradient add
        del dn2

        # ---- attention: x2 = x + o_proj(attn(rope(qkv(rmsnorm(x))))) ----
        do = dgrad(dx2, wo, meta.woT)
        if nig[5]:
            g[5] = wgrad("o", dx2, o)
        dqkv = torch.empty_like(qkv)
        if ops.FUSE_ROPE_BWD:  # inverse RoPE of dQ/dK in the backward epilogues (hd 128), else after
            ops.attn_bwd_rope(qkv[:, :qw], qkv[:, qw:qw + kw], qkv[:, qw + kw:], o, do, lse,
                              dqkv[:, :qw], dqkv[:, qw:qw + kw], dqkv[:, qw + kw:], B, H, KVH, T, hd,
                              hd ** -0.5, True, meta.cos, meta.sin, pos_ids=meta.pos_ids,
                              window=getattr(meta, "window", 0))
        else:
            ops.attn_bwd(qkv[:, :qw], qkv[:, qw:qw + kw], qkv[:, qw + kw:], o, do, lse, dqkv[:, :qw],
                         dqkv[:, qw:qw + kw], dqkv[:, qw + kw:], B, H, KVH, T, T, hd, hd ** -0.5, True,
                         window=getattr(meta, "window", 0))
            ops.rope_(dqkv, T, meta.cos, meta.sin, H + KVH, hd, inverse=True, pos_ids=meta.pos_ids)
        del do
        dn1 = dgrad(dqkv, wqkv, meta.wqkvT)
        if nig[2] or nig[3] or nig[4]:
            h1, _ = ops.rmsnorm_fwd(x, n1, eps)
            dwqkv = wgrad("qkv", dqkv, h1)
            del h1
            if dwqkv is None:
                pass
            elif ctx.split_qkv:
                g[2], g[3], g[4] = dwqkv[:qw], dwqkv[qw:qw + kw], dwqkv[qw + kw:]
            else:
                g[2] = dwqkv
        del dqkv
        if nig[1]:
            g[1] = ops.colsum(dn1, x, None, rstd1)
        if nig[0]:
            g[0] = ops.rmsnorm_bwd(dn1, x, n1, rstd1, dres=dx2)
        return tuple(g)


# =================================================================================================
class LMHeadCEFn(torch.autograd.Function):
    """lm_head + shifted next-token cross-entropy (language_model/ola_llama.py:121-136) without ever
    holding fp32 logits: per row-chunk, logits (bf16, as the reference's bf16 GEMM produces them)
    → online-softmax CE + in-place gradient → dgrad GEMM.  The hidden-state gradient is produced in
    the forward pass and only scaled by grad_output in backward."""

    @staticmethod
    def forward(ctx, hidden, w, labels, T, chunk_rows, wt=None, compact=None):
        M_full, D = hidden.shape
        V = w.shape[0]
        dev = hidden.device
        need_dh = hidden.requires_grad
        need_dw = w.requires_grad
        shift = True
        inv = None
        if compact is not None:
            # score only the rows whose shifted label is not -100 (host-built index, no sync):
            # the others add nothing to the loss or to any gradient
            rows, labels, inv, _ = compact
            hidden = ops.gather_rows(rows, [hidden], D)
            shift, T = False, 1
        M = hidden.shape[0]
        count = ops.ce_count(labels, T, shift=shift)
        row_loss = torch.empty((M,), dtype=torch.float32, device=dev)
        dh = torch.empty((M, D), dtype=BF16, device=dev) if need_dh else None
        dw = None
        logits = torch.empty((min(chunk_rows, M), V), dtype=BF16, device=dev)
        for r0 in range(0, M, chunk_rows):
            r1 = min(M, r0 + chunk_rows)
            lg = logits[: r1 - r0]
            ops.gemm(hidden[r0:r1], w, out=lg)
            ops.ce_fwd_bwd_(lg, labels, r0, T, row_loss, count, 1.0, need_dh or need_dw, shift=shift)
            if need_dh:
                if wt is not None:
                    ops.gemm(lg, wt, out=dh[r0:r1])
                else:
                    ops.gemm(lg, w, b_layout=1, out=dh[r0:r1])
            if need_dw:
                if dw is None:
                    dw = ops.gemm(lg, hidden[r0:r1], a_layout=1, b_layout=1)
                else:
                    ops.gemm(lg, hidden[r0:r1], a_layout=1, b_layout=1, residual=dw, out=dw)
        loss = ops.ce_finalize(row_loss, count)
        if inv is not None and dh is not None:
            dh = ops.gather_rows(inv, [dh], D)  # back to all rows; unscored rows get zero gradient
        ctx.save_for_backward(dh, dw)
        return loss

    @staticmethod
    def backward(ctx, gout):
        dh, dw = ctx.saved_tensors
        gout = gout.contiguous().float()
        gh = ops.scale_dev(dh, gout) if dh is not None else None
        gw = ops.scale_dev(dw, gout) if dw is not None else None
        return gh, gw, None, None, None, None, None


def lm_head_logits(hidden, w):
    """Materialise logits only when a caller asks for `.logits` (API field, fp32 in the reference)."""
    return ops.gemm(hidden, w)


# =================================================================================================
class GroupMeanFn(torch.autograd.Function):
    """Task-token pooling param[576,D].view(8,72,D).mean(1) (model/ola_arch.py:225-228)."""

    @staticmethod
    def forward(ctx, x, groups, gsize):
        ctx.groups, ctx.gsize = groups, gsize
        return ops.group_mean(x, groups, gsize)

    @staticmethod
    def backward(ctx, dy):
        return ops.group_mean_bwd(dy.contiguous(), ctx.groups, ctx.gsize), None, None


# =================================================================================================
class SpliceFn(torch.autograd.Function):
    """prepare_inputs_labels_for_multimodal's tensor work (model/ola_arch.py:345-444) as ONE gather
    driven by a host-built plan: rows come from the embedding table, the projected image features,
    or the pooled task tokens; padding rows are zero."""

    @staticmethod
    def forward(ctx, embed_w, img_feats, task_rows, plan):
        D = embed_w.shape[1]
        out = ops.gather_rows(plan.index, [embed_w, img_feats, task_rows], D, kind=plan.kind)
        ctx.plan = plan
        ctx.embed_shape = embed_w.shape
        return out

    @staticmethod
    def backward(ctx, dy):
        plan = ctx.plan
        dy = dy.contiguous()
        D = dy.shape[1]
        g_embed = g_img = g_task = None
        if ctx.needs_input_grad[1]:
            g_img = ops.gather_rows(plan.inv_img, [dy], D)
        if ctx.needs_input_grad[2] and plan.inv_task is not None:
            g_task = ops.gather_sum_rows(plan.inv_task, plan.B_cols, dy, D)
        if ctx.needs_input_grad[0]:
            acc = torch.zeros(ctx.embed_shape, dtype=torch.float32, device=dy.device)
            ops.scatter_add_rows(acc, plan.embed_scatter, dy)
            g_embed = ops.cast_bf16(acc)
        return g_embed, g_img, g_task, None


# =================================================================================================
class ResamplerFn(torch.autograd.Function):
    """TaskTokenResampler (depth=1) of the embedding-predictor heads, including the token selection
    of forward_emb_predictor:  model/multimodal_projector/resampler.py:46-75, 202-224 and
    model/language_model/base_ola_vlm.py:413-443.

    Rows are organised as [all context rows (B·nk) ; all latent rows (B·nq)] so that proj_in, to_kv
    and their gradients are single GEMMs and the Perceiver keys cat(x, latents) are two K/V segments
    of the attention kernel instead of a materialised concat.
    """

    @staticmethod
    def forward(ctx, state, special, w_in, b_in, n1w, n1b, n2w, n2b, wq, wkv, wout, fnw, fnb, wf1,
                wf2, wpo, bpo, now, nob, plan):
        B, nk, nq = plan.B, plan.nk, plan.nq
        D = state.shape[1]
        dim = w_in.shape[0]
        dev = state.device
        eps = 1e-5
        nx = B * nk
        U = torch.empty((nx + B * nq, D), dtype=BF16, device=dev)
        ops.gather_rows(plan.ctx_index, [state], D, out=U[:nx])
        if special is not None:
            ops.gather_rows(plan.lat_index, [special], D, out=U[nx:])
        else:  # gen: mean of this task's 8 token states (resampler.py:207-212 with num_queries=1)
            G = ops.gather_rows(plan.gen_index, [state], D)
            ops.group_mean(G, B, plan.nt, out=U[nx:])
            del G
        P = ops.gemm(U, w_in, bias=b_in)
        N = torch.empty_like(P)
        _, mean1, rstd1 = ops.layernorm_fwd(P[:nx], n1w, n1b, eps, out=N[:nx])
        _, mean2, rstd2 = ops.layernorm_fwd(P[nx:], n2w, n2b, eps, out=N[nx:])
        q = ops.gemm(N[nx:], wq)
        kv = ops.gemm(N, wkv)
        inner = wq.shape[0]
        heads = plan.heads
        hd = inner // heads
        ao, lse = ops.attn_fwd(q, kv[:nx, :inner], kv[:nx, inner:], B, heads, heads, nq, nk, hd,
                               hd ** -0.5, False, k2=kv[nx:, :inner], v2=kv[nx:, inner:], sk2=nq)
        L1 = ops.gemm(ao, wout, residual=P[nx:])
        F0, meanf, rstdf = ops.layernorm_fwd(L1, fnw, fnb, eps)
        F1, F1pre = ops.gemm(F0, wf1, act=ACT_GELU, want_pre=True)
        L2 = ops.gemm(F1, wf2, residual=L1)
        Y = ops.gemm(L2, wpo, bias=bpo)
        E, meano, rstdo = ops.layernorm_fwd(Y, now, nob, eps)
        ctx.plan = plan
        ctx.has_special = special is not None
        ctx.state_shape = state.shape
        ctx.save_for_backward(U, w_in, n1w, n2w, wq, wkv, wout, fnw, wf1, wf2, wpo, now, P, N, mean1,
                              rstd1, mean2, rstd2, q, kv, ao, lse, L1, meanf, rstdf, F0, F1, F1pre, L2,
                              Y, meano, rstdo)
        return E

    @staticmethod
    def backward(ctx, dE):
        (U, w_in, n1w, n2w, wq, wkv, wout, fnw, wf1, wf2, wpo, now, P, N, mean1, rstd1, mean2, rstd2,
         q, kv, ao, lse, L1, meanf, rstdf, F0, F1, F1pre, L2, Y, meano, rstdo) = ctx.saved_tensors
        plan = ctx.plan
        B, nk, nq = plan.B, plan.nk, plan.nq
        nx = B * nk
        D = U.shape[1]
        inner = wq.shape[0]
        heads = plan.heads
        hd = inner // heads
        dE = dE.contiguous()
        g = [None] * 20
        f32 = torch.float32

        # norm_out, proj_out
        g[17] = ops.colsum(dE, Y, meano, rstdo)
        g[18] = ops.colsum(dE)
        dY = ops.layernorm_bwd(dE, Y, now, meano, rstdo)
        g[15] = ops.gemm(dY, L2, a_layout=1, b_layout=1)
        g[16] = ops.colsum(dY)
        dL2 = ops.gemm(dY, wpo, b_layout=1)
        # feed-forward: L2 = L1 + W2·gelu(W1·LN(L1))
        g[14] = ops.gemm(dL2, F1, a_layout=1, b_layout=1)
        dF1 = ops.act_bwd(F1pre, ops.gemm(dL2, wf2, b_layout=1), ACT_GELU)
        g[13] = ops.gemm(dF1, F0, a_layout=1, b_layout=1)
        dF0 = ops.gemm(dF1, wf1, b_layout=1)
        g[11] = ops.colsum(dF0, L1, meanf, rstdf)
        g[12] = ops.colsum(dF0)
        dL1 = ops.layernorm_bwd(dF0, L1, fnw, meanf, rstdf, dres=dL2)
        # attention: L1 = Pl + Wout·attn(q, kv)
        g[10] = ops.gemm(dL1, ao, a_layout=1, b_layout=1)
        dao = ops.gemm(dL1, wout, b_layout=1)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        ops.attn_bwd(q, kv[:nx, :inner], kv[:nx, inner:], ao, dao, lse, dq, dkv[:nx, :inner],
                     dkv[:nx, inner:], B, heads, heads, nq, nk, hd, hd ** -0.5, False,
                     k2=kv[nx:, :inner], v2=kv[nx:, inner:], sk2=nq, dk2=dkv[nx:, :inner],
                     dv2=dkv[nx:, inner:])
        g[8] = ops.gemm(dq, N[nx:], a_layout=1, b_layout=1)
        g[9] = ops.gemm(dkv, N, a_layout=1, b_layout=1)
        dN = ops.gemm(dkv, wkv, b_layout=1)
        ops.gemm(dq, wq, b_layout=1, residual=dN[nx:], out=dN[nx:])  # latents feed both q and kv
        # the two input LayerNorms (+ residual path of the latents)
        g[4] = ops.colsum(dN[:nx], P[:nx], mean1, rstd1)
        g[5] = ops.colsum(dN[:nx])
        g[6] = ops.colsum(dN[nx:], P[nx:], mean2, rstd2)
        g[7] = ops.colsum(dN[nx:])
        dP = torch.empty_like(P)
        ops.layernorm_bwd(dN[:nx], P[:nx], n1w, mean1, rstd1, out=dP[:nx])
        ops.layernorm_bwd(dN[nx:], P[nx:], n2w, mean2, rstd2, dres=dL1, out=dP[nx:])
        # proj_in (shared by context and latents)
        g[2] = ops.gemm(dP, U, a_layout=1, b_layout=1)
        g[3] = ops.colsum(dP)
        dU = ops.gemm(dP, w_in, b_layout=1)
        # scatter back to the layer state / the task-token parameter
        if ctx.needs_input_grad[0]:
            dstate = ops.gather_rows(plan.inv_ctx, [dU], D)
            if not ctx.has_special:
                dG = ops.group_mean_bwd(dU[nx:], B, plan.nt)
                ops.add_rows_(dstate, plan.gen_index, dG)
            g[0] = dstate
        if ctx.has_special and ctx.needs_input_grad[1]:
            g[1] = ops.gather_sum_rows(plan.inv_lat, B, dU[nx:], D)
        return tuple(g)


# =================================================================================================
class DistillLossFn(torch.autograd.Function):
    """_emb_loss (language_model/base_ola_vlm.py:289-320) + calculate_contrastive_loss
    (ola_utils.py:108-125): smooth-L1 + 0.3·InfoNCE over flattened L2-normalised embeddings against
    (all-gathered) targets, fp32 math in one fused reduction. Returns [loss, sl1, contrastive]."""

    @staticmethod
    def forward(ctx, pred, tgt_all, tau, mask, off, cw):
        out4, coef, _ = ops.distill_loss_fwd(pred, tgt_all, off, tau, mask, cw)
        ctx.off = off
        ctx.save_for_backward(pred, tgt_all, coef, out4)
        return out4[:3].clone()

    @staticmethod
    def backward(ctx, gout):
        pred, tgt_all, coef, out4 = ctx.saved_tensors
        # only the total (element 0) carries gradient; sl1/contrastive are reporting outputs
        g0 = gout[0:1].contiguous().float()
        dpred = ops.distill_loss_bwd(pred, tgt_all, ctx.off, coef, g0) if ctx.needs_input_grad[0] else None
        dtau = (out4[3] * g0[0]).reshape(1) if ctx.needs_input_grad[2] else None
        return dpred, None, dtau, None, None, None
