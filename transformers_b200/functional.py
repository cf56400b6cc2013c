"""torch.autograd.Function wrappers: each forward / backward is a short sequence of C-ABI kernel launches.

Functions are re-entrant (no hidden global state) so the reference's GradientCheckpointingLayer
(modeling_layers.py:49-110) can replay them.
"""
from __future__ import annotations

import torch

from . import ops


def _needs(ctx, i):
    return ctx.needs_input_grad[i]


def _tp_reduce_async(t, group):
    """Start an all-reduce(sum) of ``t`` on the communicator's stream; returns the work handle (None without a group)."""
    if group is None:
        return None
    import torch.distributed as dist

    return dist.all_reduce(t, group=group, async_op=True)


def _tp_unpack(tp):
    """tp descriptor -> (group, mode, sequence-parallel state).  Modes: "col" / "row" (replicated activations, all-reduce)
    and "col_sp" / "row_sp" (token-sharded activations between the blocks, parallel.SequenceParallelState)."""
    if tp is None:
        return None, None, None
    return tp[0], tp[1], (tp[2] if len(tp) > 2 else None)


def _peer_reduce_scatter(part, st):
    """Reduce-scatter of the partial sums every rank just wrote into its peer-mapped buffer ``part`` [T, C]: barrier, then
    one kernel pulls this rank's rows from all peers over NVLink and sums them (csrc/peer.cu).  Returns [T/N, C]."""
    T, C = part.shape
    rows = T // st.world
    st.peer.publish_partial()
    out = torch.empty(rows, C, device=part.device, dtype=part.dtype)
    return ops.pull_reduce(st.peer.partial_ptrs(), st.rank * rows * C, rows * C, out)


def _peer_gemm_reduce_scatter(x2, w, st, b_mn=False):
    """Rowwise GEMM + reduce-scatter over peer memory.  With the scatter epilogue the GEMM itself delivers every finished
    tile to its owner over NVLink (one kernel for the GEMM and the transfer) and the reduction reads local slots; otherwise
    the partial stays local and the owners pull it.  Returns this rank's [T/N, C] rows of the summed result."""
    ws = st.peer
    T = x2.shape[0]
    C = w.shape[1] if b_mn else w.shape[0]
    rows = T // st.world
    if ws.scatter_epilogue and rows % 256 == 0:
        dest, mine = ws.next_staging(rows, C)
        ops.gemm_scatter(x2, w, dest, st.rank, b_mn=b_mn)
        ws.publish_partial()
        out = torch.empty(rows, C, device=x2.device, dtype=x2.dtype)
        return ops.pull_reduce(mine, 0, rows * C, out)
    part = ws.next_partial(T, C)
    ops.gemm(x2, w, b_mn=b_mn, out=part)
    return _peer_reduce_scatter(part, st)


def _peer_all_gather(local2, st):
    """All-gather of token shards over peer memory: publish the shard, barrier, then copy every peer's shard out with the
    copy engines on a side stream (no SM time); the own rows are a local copy.  Returns [T, K]."""
    rows, K = local2.shape
    ws = st.peer
    ws.publish_shard(local2)
    full = torch.empty(rows * st.world, K, device=local2.device, dtype=local2.dtype)
    full[st.rank * rows:(st.rank + 1) * rows].copy_(local2)
    for i in range(1, st.world):
        src = (st.rank + i) % st.world
        with ws.copy_context(i):  # pulls from different peers on different copy engines, in parallel
            full[src * rows:(src + 1) * rows].copy_(ws.peer_shard(src, rows, K), non_blocking=True)
    ws.join_copies()
    return full


def _sp_gather_gemm(x_local2, w, st, glu=None):
    """Colwise block entry under sequence parallelism: y = all_gather(x) @ w^T with the all-gather of chunk c+1 running
    on the communicator's stream while chunk c is in the GEMM.  Returns (x_full [T,K] -- kept for the wgrad --, y [T,N]);
    with ``glu`` (the GeGLU flag; w is then the block-interleaved gate|up weight) the GEMM runs the gated activation in its
    epilogue and the result is (x_full, y, h [T, N/2])."""
    from .parallel import sp_all_gather

    if st.peer is not None:
        x_full = _peer_all_gather(x_local2, st)
        if glu is not None:
            return (x_full, *ops.gemm_glu(x_full, w, glu))
        return x_full, ops.gemm(x_full, w)
    x_full, works = sp_all_gather(x_local2, st)
    y = x_full.new_empty(x_full.shape[0], w.shape[0])
    h = x_full.new_empty(x_full.shape[0], w.shape[0] // 2) if glu is not None else None
    for rows, work in zip(st.chunk_rows(x_full.shape[0]), works):
        work.wait()
        if glu is not None:
            ops.gemm_glu(x_full[rows], w, glu, gu_out=y[rows], h_out=h[rows])
        else:
            ops.gemm(x_full[rows], w, out=y[rows])
    return (x_full, y, h) if glu is not None else (x_full, y)


def _sp_dgrad_scatter(dy2, w, st):
    """Colwise block entry, backward: dX = dY @ W is a partial sum over the ranks' column shards -> reduce-scatter it to
    the token shards, chunk c's reduce-scatter running under chunk c+1's GEMM (and under the caller's wgrad GEMM).
    Returns (dx_local [T/N,K], pending works, buffers to keep alive until the works are waited on)."""
    from .parallel import sp_reduce_scatter_chunk

    T = dy2.shape[0]
    if st.peer is not None:
        return _peer_gemm_reduce_scatter(dy2, w, st, b_mn=True), [], None
    dx_local = dy2.new_empty(T // st.world, w.shape[1])
    works, keep = [], []
    for c, rows in enumerate(st.chunk_rows(T)):
        part = ops.gemm(dy2[rows], w, b_mn=True)
        keep.append(part)
        works.append(sp_reduce_scatter_chunk(part, dx_local, c, st))
    return dx_local, works, keep


def _sp_gemm_scatter(x2, w, st):
    """Rowwise block exit under sequence parallelism: reduce_scatter(x @ w^T), chunk c's reduce-scatter under chunk c+1's
    GEMM.  Returns y_local [T/N, N]."""
    from .parallel import sp_reduce_scatter_chunk

    T = x2.shape[0]
    if st.peer is not None:
        return _peer_gemm_reduce_scatter(x2, w, st)
    y_local = x2.new_empty(T // st.world, w.shape[0])
    works, keep = [], []
    for c, rows in enumerate(st.chunk_rows(T)):
        part = ops.gemm(x2[rows], w)
        keep.append(part)
        works.append(sp_reduce_scatter_chunk(part, y_local, c, st))
    for work in works:
        work.wait()
    return y_local


def _sp_gather_dgrad(dy_local2, w, st):
    """Rowwise block exit, backward: dY arrives token-sharded -> all-gather it (chunk c+1 under chunk c's dgrad GEMM).
    Returns (dy_full [T,N] -- for the wgrad --, dx [T,K])."""
    from .parallel import sp_all_gather

    if st.peer is not None:
        dy_full = _peer_all_gather(dy_local2, st)
        return dy_full, ops.gemm(dy_full, w, b_mn=True)
    dy_full, works = sp_all_gather(dy_local2, st)
    dx = dy_full.new_empty(dy_full.shape[0], w.shape[1])
    for rows, work in zip(st.chunk_rows(dy_full.shape[0]), works):
        work.wait()
        ops.gemm(dy_full[rows], w, b_mn=True, out=dx[rows])
    return dy_full, dx


class FusedLinearFn(torch.autograd.Function):
    """y = x @ cat(weights)^T  -- one GEMM for several nn.Linear layers that share their input
    (q/k/v: models/llama/modeling_llama.py:254-256; gate/up: :174-176) or a single one (o_proj, down_proj, lm_head).

    ``w_fused`` is the row-wise concatenation [sum(N_i), K] (maintained by the calling module); ``weights`` are the
    individual parameters, passed so autograd routes their gradients; gradients are row-slices of one fused wgrad.

    Tensor parallelism (``tp = (group, mode)``), collectives overlapped with our own GEMMs instead of the reference's
    blocking DTensor redistributes (distributed/tensor_parallel.py:219-226, :320-328):
      mode "col": the input gradient is a partial sum -> its all-reduce runs on the NCCL stream while the wgrad GEMM runs;
      mode "row": the output is a partial sum -> the GEMM is issued in two row halves, the first half's all-reduce
                  overlaps the second half's GEMM."""

    @staticmethod
    def forward(ctx, x, w_fused, tp, *weights):
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        N = w_fused.shape[0]
        group, mode, st = _tp_unpack(tp)
        ctx.splits = [w.shape[0] for w in weights]
        ctx.x_shape = x.shape
        ctx.tp = tp
        if mode == "col_sp":  # x is this rank's token shard [1, T/N, K] -> y [1, T, N]
            x_full, y2 = _sp_gather_gemm(x2, w_fused, st)
            ctx.save_for_backward(x_full, w_fused)
            return y2.view(1, -1, N)
        if mode == "row_sp":  # x holds all tokens [.., K/N] -> y is this rank's token shard [1, T/N, N]
            ctx.save_for_backward(x2, w_fused)
            return _sp_gemm_scatter(x2, w_fused, st).view(1, -1, N)
        y = torch.empty(*x.shape[:-1], N, device=x.device, dtype=x.dtype)
        y2 = y.view(-1, N)
        T = x2.shape[0]
        if group is not None and mode == "row" and T >= 512:
            h = (T // 2 + 127) // 128 * 128
            ops.gemm(x2[:h], w_fused, out=y2[:h])
            w1 = _tp_reduce_async(y2[:h], group)
            ops.gemm(x2[h:], w_fused, out=y2[h:])
            w2 = _tp_reduce_async(y2[h:], group)
            w1.wait()
            w2.wait()
        else:
            ops.gemm(x2, w_fused, out=y2)
            if group is not None and mode == "row":
                _tp_reduce_async(y2, group).wait()
        ctx.save_for_backward(x2, w_fused)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, w_fused = ctx.saved_tensors
        N = w_fused.shape[0]
        dy2 = dy.reshape(-1, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        group, mode, st = _tp_unpack(ctx.tp)
        dx, work, works, keep = None, None, (), None
        if mode == "row_sp":
            dy2, dx = _sp_gather_dgrad(dy2, w_fused, st)  # dy2 is now the gathered [T, N]
            dx = dx.view(ctx.x_shape)
        elif mode == "col_sp":
            if _needs(ctx, 0):
                dx, works, keep = _sp_dgrad_scatter(dy2, w_fused, st)  # overlaps the wgrad GEMM below
                dx = dx.view(ctx.x_shape)
        elif _needs(ctx, 0):
            dx = ops.gemm(dy2, w_fused, b_mn=True).view(ctx.x_shape)  # dX = dY W : B stored [K'=N, N'=K]
            if mode == "col":
                work = _tp_reduce_async(dx, group)  # overlaps the wgrad GEMM below
        grads_w = [None] * len(ctx.splits)
        if any(ctx.needs_input_grad[3:]):
            dw = ops.gemm(dy2, x2, a_mn=True, b_mn=True)  # dW[N,K] = dY^T X
            off = 0
            for i, n in enumerate(ctx.splits):
                if ctx.needs_input_grad[3 + i]:
                    grads_w[i] = dw[off:off + n]
                off += n
        if work is not None:
            work.wait()
        for w_ in works:
            w_.wait()
        del keep
        return (dx, None, None, *grads_w)


class GateUpGluFn(torch.autograd.Function):
    """h = act(x Wg^T) * (x Wu^T) -- the first two thirds of LlamaMLP.forward (models/llama/modeling_llama.py:174-176) as ONE
    kernel: the gate|up GEMM on the block-interleaved weight with the gated activation in its epilogue (csrc/gemm2.cu GLU
    mode); the projections are still written (interleaved) because the backward needs them, but never read back in the
    forward.  Backward: d(gate|up) from the GLU backward kernel on the interleaved layout, then the usual dgrad / wgrad GEMMs
    on the interleaved weight; the weight gradient is de-interleaved into the gate_proj / up_proj gradients.
    ``w_ilv`` is maintained by the calling module (modules.interleaved_weight); tp modes as in FusedLinearFn ("col" /
    "col_sp"): the block-interleaving is applied to the rank's own column shard."""

    @staticmethod
    def forward(ctx, x, w_ilv, gelu, tp, wg, wu):
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        _, mode, st = _tp_unpack(tp)
        I = w_ilv.shape[0] // 2
        if mode == "col_sp":
            x2, gu, h = _sp_gather_gemm(x2, w_ilv, st, glu=gelu)
            out_shape = (1, x2.shape[0], I)
        else:
            gu, h = ops.gemm_glu(x2, w_ilv, gelu)
            out_shape = (*x.shape[:-1], I)
        ctx.save_for_backward(x2, w_ilv, gu)
        ctx.cfg = (gelu, tp, x.shape)
        return h.view(out_shape)

    @staticmethod
    def backward(ctx, dh):
        x2, w_ilv, gu = ctx.saved_tensors
        gelu, tp, x_shape = ctx.cfg
        group, mode, st = _tp_unpack(tp)
        dgu = ops.glu_bwd(dh.reshape(-1, dh.shape[-1]), gu, gelu, interleaved=True)
        dx, work, works, keep = None, None, (), None
        if _needs(ctx, 0):
            if mode == "col_sp":
                dx, works, keep = _sp_dgrad_scatter(dgu, w_ilv, st)
            else:
                dx = ops.gemm(dgu, w_ilv, b_mn=True)
                if mode == "col":
                    work = _tp_reduce_async(dx, group)  # overlaps the wgrad GEMM below
            dx = dx.view(x_shape)
        dwg = dwu = None
        if _needs(ctx, 4) or _needs(ctx, 5):
            dw = ops.gemm(dgu, x2, a_mn=True, b_mn=True)
            dwg, dwu = ops.deinterleave_gate_up(dw)
        if work is not None:
            work.wait()
        for w_ in works:
            w_.wait()
        del keep
        return dx, None, None, None, (dwg if _needs(ctx, 4) else None), (dwu if _needs(ctx, 5) else None)


class RMSNormFn(torch.autograd.Function):
    """LlamaRMSNorm / Gemma2RMSNorm (models/llama/modeling_llama.py:62-67, models/gemma2/modeling_gemma2.py:55-63)."""

    @staticmethod
    def forward(ctx, x, weight, eps, gemma):
        y, rstd, _ = ops.rmsnorm_fwd(x, weight, eps, gemma)
        ctx.save_for_backward(x, weight, rstd)
        ctx.gemma = gemma
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, rstd = ctx.saved_tensors
        dx, dw = ops.rmsnorm_bwd(dy, x, weight, rstd, ctx.gemma)
        return dx if _needs(ctx, 0) else None, dw if _needs(ctx, 1) else None, None, None


class AddRMSNormFn(torch.autograd.Function):
    """r = x + residual; y = RMSNorm(r) in ONE pass over the row (LlamaDecoderLayer.forward models/llama/modeling_llama.py:
    317-321 does `residual + hidden_states` and `post_attention_layernorm` as two).  Returns (r, y)."""

    @staticmethod
    def forward(ctx, x, residual, weight, eps, gemma):
        y, rstd, r = ops.rmsnorm_fwd(x, weight, eps, gemma, residual=residual)
        ctx.save_for_backward(r, weight, rstd)
        ctx.gemma = gemma
        return r, y

    @staticmethod
    def backward(ctx, dr, dy):
        r, weight, rstd = ctx.saved_tensors
        dxn, dw = ops.rmsnorm_bwd(dy.contiguous(), r, weight, rstd, ctx.gemma)
        d = ops.add(dr, dxn) if dr is not None else dxn  # both addends of r receive the same gradient
        return d, d, (dw if _needs(ctx, 2) else None), None, None


class AddFn(torch.autograd.Function):
    """a + b on our kernel (the second residual add of the decoder layer, :323)."""

    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


class GluFn(torch.autograd.Function):
    """act(gate) * up on the packed [.., 2I] projection output (LlamaMLP.forward :174-176)."""

    @staticmethod
    def forward(ctx, gu, gelu):
        ctx.save_for_backward(gu)
        ctx.gelu = gelu
        return ops.glu_fwd(gu, gelu)

    @staticmethod
    def backward(ctx, dh):
        (gu,) = ctx.saved_tensors
        return ops.glu_bwd(dh, gu, ctx.gelu), None


def _attn_fwd_any(q, k, v, scale, causal, window, softcap, kv_start, kv_end, segments):
    """Attention forward over [B, S, h, D] views.  ``segments`` (packed batch: [[(start, end), ...] per row]) runs the
    kernel once per sequence on row slices of the same buffers -- varlen attention without un-padding copies, the way
    the reference's flash path calls flash_attn_varlen_func (modeling_flash_attention_utils.py:796-822).
    Returns (out, [lse, ...])."""
    if segments is None:
        out, lse = ops.attn_fwd(q, k, v, scale=scale, causal=causal, window=window, softcap=softcap, kv_start=kv_start,
                                kv_end=kv_end)
        return out, [lse]
    if q.shape[1] != k.shape[1]:
        raise ops.B200Error("b200 attention: packed batches need q_len == kv_len (no KV cache)")
    out = torch.empty(q.shape, device=q.device, dtype=q.dtype)
    lses = []
    for b, rows in enumerate(segments):
        for s, e in rows:
            _, lse = ops.attn_fwd(q[b:b + 1, s:e], k[b:b + 1, s:e], v[b:b + 1, s:e], scale=scale, causal=causal, window=window,
                                  softcap=softcap, out=out[b:b + 1, s:e])
            lses.append(lse)
    return out, lses


def _attn_bwd_any(q, k, v, out, dout, lses, dq, dk, dv, scale, causal, window, softcap, kv_start, kv_end, segments):
    if segments is None:
        ops.attn_bwd(q, k, v, out, dout, lses[0], dq, dk, dv, scale=scale, causal=causal, window=window, softcap=softcap,
                     kv_start=kv_start, kv_end=kv_end)
        return
    i = 0
    for b, rows in enumerate(segments):
        for s, e in rows:
            sl = (slice(b, b + 1), slice(s, e))
            ops.attn_bwd(q[sl], k[sl], v[sl], out[sl], dout[sl], lses[i], dq[sl], dk[sl], dv[sl], scale=scale, causal=causal,
                         window=window, softcap=softcap)
            i += 1


class QKVRopeAttentionFn(torch.autograd.Function):
    """Fused attention block up to (not including) o_proj, LlamaAttention.forward models/llama/modeling_llama.py:251-279:
    packed QKV projection (one GEMM) -> RoPE in place on the private projection buffer (apply_rotary_pos_emb :138-160)
    -> flash attention reading q/k/v as strided views of that buffer (:191-213).  Returns [B, S, Hq*D].
    Backward: attention bwd writes dq|dk|dv straight into one packed buffer -> RoPE^T in place -> dgrad + wgrad GEMMs."""

    @staticmethod
    def forward(ctx, x, w_fused, cos, sin, cfg, kv_start, kv_end, tp, segments, *weights):
        Hq, Hkv, D, scale, causal, window, softcap = cfg
        ctx.tp = tp
        ctx.segments = segments
        _, mode, st = _tp_unpack(tp)
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        if mode == "col_sp":  # x is this rank's token shard; (B, S) are those of the whole batch
            B, S = st.full_shape[:2]
            x2, qkv = _sp_gather_gemm(x2, w_fused, st)
            qkv = qkv.view(B, S, -1)
        else:
            B, S = x.shape[:2]
            qkv = ops.gemm(x2, w_fused).view(B, S, -1)
        ops.rope_(qkv, cos, sin, Hq + Hkv, D)
        q = qkv[..., : Hq * D].view(B, S, Hq, D)
        k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D)
        v = qkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
        out, lses = _attn_fwd_any(q, k, v, scale, causal, window, softcap, kv_start, kv_end, segments)
        ctx.save_for_backward(x2, w_fused, qkv, out, cos, sin, kv_start, kv_end, *lses)
        ctx.cfg = cfg
        ctx.splits = [w.shape[0] for w in weights]
        ctx.x_shape = x.shape
        return out.view(B, S, Hq * D)

    @staticmethod
    def backward(ctx, dout):
        x2, w_fused, qkv, out, cos, sin, kv_start, kv_end, *lses = ctx.saved_tensors
        Hq, Hkv, D, scale, causal, window, softcap = ctx.cfg
        B, S, W = qkv.shape
        q = qkv[..., : Hq * D].view(B, S, Hq, D)
        k = qkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D)
        v = qkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
        dout4 = dout.reshape(B, S, Hq, D)
        if not dout4.is_contiguous():
            dout4 = dout4.contiguous()
        dqkv = torch.empty_like(qkv)
        dq = dqkv[..., : Hq * D].view(B, S, Hq, D)
        dk = dqkv[..., Hq * D:(Hq + Hkv) * D].view(B, S, Hkv, D)
        dv = dqkv[..., (Hq + Hkv) * D:].view(B, S, Hkv, D)
        _attn_bwd_any(q, k, v, out, dout4, lses, dq, dk, dv, scale, causal, window, softcap, kv_start, kv_end, ctx.segments)
        ops.rope_(dqkv, cos, sin, Hq + Hkv, D, backward=True)
        d2 = dqkv.view(B * S, W)
        _, mode, st = _tp_unpack(ctx.tp)
        dx, work, works, keep = None, None, (), None
        if mode == "col_sp":
            if _needs(ctx, 0):
                dx, works, keep = _sp_dgrad_scatter(d2, w_fused, st)  # reduce-scatters overlap the wgrad
                dx = dx.view(ctx.x_shape)
        elif _needs(ctx, 0):
            dx = ops.gemm(d2, w_fused, b_mn=True).view(ctx.x_shape)
            work = _tp_reduce_async(dx, ctx.tp[0]) if ctx.tp is not None else None  # overlaps the wgrad
        grads_w = [None] * len(ctx.splits)
        if any(ctx.needs_input_grad[9:]):
            dw = ops.gemm(d2, x2, a_mn=True, b_mn=True)
            off = 0
            for i, n in enumerate(ctx.splits):
                if ctx.needs_input_grad[9 + i]:
                    grads_w[i] = dw[off:off + n]
                off += n
        if work is not None:
            work.wait()
        for w_ in works:
            w_.wait()
        del keep
        return (dx, None, None, None, None, None, None, None, None, *grads_w)


class FlashAttentionFn(torch.autograd.Function):
    """Attention core on separate strided q/k/v views ([B, S, h, D]); used by the registry entry point
    (AttentionInterface signature, docs/source/en/attention_interface.md:164-175) incl. the KV-cache / decode path."""

    @staticmethod
    def forward(ctx, q, k, v, scale, causal, window, softcap, kv_start, kv_end, segments=None):
        out, lses = _attn_fwd_any(q, k, v, scale, causal, window, softcap, kv_start, kv_end, segments)
        ctx.save_for_backward(q, k, v, out, kv_start, kv_end, *lses)
        ctx.cfg = (scale, causal, window, softcap, segments)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, kv_start, kv_end, *lses = ctx.saved_tensors
        scale, causal, window, softcap, segments = ctx.cfg
        dout = dout.contiguous()
        dq = torch.empty(q.shape, device=q.device, dtype=q.dtype)
        dk = torch.empty(k.shape, device=k.device, dtype=k.dtype)
        dv = torch.empty(v.shape, device=v.device, dtype=v.dtype)
        _attn_bwd_any(q, k, v, out, dout, lses, dq, dk, dv, scale, causal, window, softcap, kv_start, kv_end, segments)
        return dq, dk, dv, None, None, None, None, None, None, None


class QKVRopeFn(torch.autograd.Function):
    """Packed QKV projection + RoPE, returning the rotated buffer [B, S, (Hq + 2 Hkv) D] (KV-cache path: k / v then go
    through Cache.update, models/llama/modeling_llama.py:262, and attention runs through the registry entry)."""

    @staticmethod
    def forward(ctx, x, w_fused, cos, sin, n_rot, D, tp, *weights):
        ctx.tp = tp
        B, S, K = x.shape
        x2 = x.reshape(B * S, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        qkv = ops.gemm(x2, w_fused).view(B, S, -1)
        ops.rope_(qkv, cos, sin, n_rot, D)
        ctx.save_for_backward(x2, w_fused, cos, sin)
        ctx.cfg = (n_rot, D)
        ctx.splits = [w.shape[0] for w in weights]
        ctx.x_shape = x.shape
        return qkv

    @staticmethod
    def backward(ctx, dqkv):
        x2, w_fused, cos, sin = ctx.saved_tensors
        B, S, _ = ctx.x_shape
        dqkv = dqkv.contiguous().clone()
        ops.rope_(dqkv, cos, sin, *ctx.cfg, backward=True)
        d2 = dqkv.view(B * S, -1)
        dx = ops.gemm(d2, w_fused, b_mn=True).view(ctx.x_shape) if _needs(ctx, 0) else None
        work = _tp_reduce_async(dx, ctx.tp[0]) if (dx is not None and ctx.tp is not None) else None
        grads_w = [None] * len(ctx.splits)
        if any(ctx.needs_input_grad[7:]):
            dw = ops.gemm(d2, x2, a_mn=True, b_mn=True)
            off = 0
            for i, n in enumerate(ctx.splits):
                if ctx.needs_input_grad[7 + i]:
                    grads_w[i] = dw[off:off + n]
                off += n
        if work is not None:
            work.wait()
        return (dx, None, None, None, None, None, None, *grads_w)


class EmbeddingFn(torch.autograd.Function):
    """nn.Embedding gather (models/llama/modeling_llama.py:381), optional Gemma scale; backward = scatter-add."""

    @staticmethod
    def forward(ctx, ids, weight, padding_idx, scale):
        ctx.save_for_backward(ids)
        ctx.cfg = (weight.shape[0], padding_idx, scale)
        return ops.embedding_fwd(ids, weight, scale)

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        V, padding_idx, scale = ctx.cfg
        return None, ops.embedding_bwd(ids, dout, V, padding_idx, scale), None, None


class CausalLMLossFn(torch.autograd.Function):
    """ForCausalLMLoss (loss/loss_utils.py:48-70): shift labels, fp32 log-softmax, mean over valid targets."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index, num_items, shift):
        loss, lse, denom = ops.ce_fwd(logits, labels, shift, ignore_index, num_items)
        ctx.save_for_backward(logits, labels, lse, denom)
        ctx.cfg = (ignore_index, shift)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, lse, denom = ctx.saved_tensors
        ignore_index, shift = ctx.cfg
        return ops.ce_bwd(logits, labels, lse, dloss, denom, shift, ignore_index), None, None, None, None


class FusedHeadLossFn(torch.autograd.Function):
    """lm_head + ForCausalLMLoss (models/llama/modeling_llama.py:479-484, loss/loss_utils.py:48-70) without ever holding
    the [T, V] logits: the T = B*S tokens are walked in row chunks; per chunk one GEMM produces that chunk's logits into a
    re-used buffer, the CE kernels turn them into per-row losses and -- when gradients are wanted -- straight into
    d(logits) (the 1/denominator is known up front: it only depends on the labels), which the dgrad GEMM consumes for
    d(hidden) and the wgrad GEMM accumulates into d(W) through the TMA reduce-add epilogue.  Same FLOPs as the unfused
    path (nothing is recomputed), -8.4 GB of live logits / d(logits) at Llama-3-8B / T = 16384 (2 x 0.5 GB buffers instead);
    ``backward`` only scales the stored gradients by the incoming scalar."""

    CHUNK_ROWS = 2048

    @staticmethod
    def forward(ctx, h, w, labels, ignore_index, num_items, shift, w_param):
        H = h.shape[-1]
        h2 = h.reshape(-1, H)
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        T, V = h2.shape[0], w.shape[0]
        labels = labels.reshape(h.shape[0], -1).to(torch.int64)
        if shift:
            tgt = torch.full_like(labels, ignore_index)
            tgt[:, :-1] = labels[:, 1:]
        else:
            tgt = labels
        tgt = tgt.reshape(T).contiguous()
        if num_items:
            denom = torch.full((1,), float(num_items), device=h.device, dtype=torch.float32)
        else:
            denom = (tgt != ignore_index).sum().to(torch.float32).reshape(1)
        want_dh, want_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[6]
        grads = want_dh or want_dw
        chunk = min(T, FusedHeadLossFn.CHUNK_ROWS)
        buf = torch.empty(chunk, V, device=h.device, dtype=h.dtype)
        dbuf = torch.empty(chunk, V, device=h.device, dtype=h.dtype) if grads else None
        dh = torch.empty_like(h2) if want_dh else None
        dw = torch.empty_like(w) if want_dw else None
        one = torch.ones(1, device=h.device, dtype=torch.float32)
        total = torch.zeros((), device=h.device, dtype=torch.float32)
        for r0 in range(0, T, chunk):
            n = min(chunk, T - r0)
            rows = slice(r0, r0 + n)
            logits = ops.gemm(h2[rows], w, out=buf[:n]).view(1, n, V)
            part, lse, _ = ops.ce_fwd(logits, tgt[rows].view(1, n), False, ignore_index, 1.0)  # num_items = 1: the plain sum
            total = total + part
            if grads:
                dl = ops.ce_bwd(logits, tgt[rows].view(1, n), lse, one, denom, False, ignore_index, out=dbuf[:n]).view(n, V)
                if want_dh:
                    ops.gemm(dl, w, b_mn=True, out=dh[rows])
                if want_dw:
                    ops.gemm(dl, h2[rows], a_mn=True, b_mn=True, out=dw, accumulate=r0 > 0)
        ctx.save_for_backward(dh, dw)
        ctx.h_shape = h.shape
        return total / denom[0]

    @staticmethod
    def backward(ctx, g):
        dh, dw = ctx.saved_tensors
        g = g.to(torch.float32)
        gh = (dh * g).to(dh.dtype).view(ctx.h_shape) if dh is not None else None
        gw = (dw * g).to(dw.dtype) if dw is not None else None
        return gh, None, None, None, None, None, gw


class VocabParallelLossFn(torch.autograd.Function):
    """ForCausalLMLoss (loss/loss_utils.py:48-70) on a vocabulary-sharded lm_head output: every rank holds the columns
    [rank*V/N, (rank+1)*V/N) of the logits.  Instead of all-gathering the logits (4.2 GB for Llama-3-8B at T = 16384) the
    ranks exchange two fp32 numbers per row: the local log-sum-exp and the target's logit (owned by exactly one rank)."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index, num_items, shift, group):
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B, S, Vl = logits.shape
        T = B * S
        lg = logits.reshape(T, Vl)
        if not lg.is_contiguous():
            lg = lg.contiguous()
        labels = labels.reshape(B, S).to(torch.int64)
        if shift:
            tgt = torch.full_like(labels, ignore_index)
            tgt[:, :-1] = labels[:, 1:]
        else:
            tgt = labels
        tgt = tgt.reshape(T)
        valid = tgt != ignore_index
        local = tgt - rank * Vl
        mine = valid & (local >= 0) & (local < Vl)
        local = torch.where(mine, local, torch.full_like(local, -1))
        lse_local = ops.ce_row_lse(lg)
        picked = lg.gather(1, local.clamp(min=0).unsqueeze(1)).squeeze(1).float()
        stats = torch.stack([lse_local, torch.where(mine, picked, torch.zeros_like(picked))])  # [2, T]
        gathered = stats.new_empty(world * 2, T)
        dist.all_gather_into_tensor(gathered, stats, group=group)
        gathered = gathered.view(world, 2, T)
        lse_global = torch.logsumexp(gathered[:, 0], dim=0)
        target_logit = gathered[:, 1].sum(dim=0)
        denom = valid.sum().float() if not num_items else torch.tensor(float(num_items), device=lg.device)
        loss = ((lse_global - target_logit) * valid).sum() / denom
        ctx.save_for_backward(lg, local, lse_global, valid, denom)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lg, local, lse_global, valid, denom = ctx.saved_tensors
        row_scale = valid.float() * (dloss.float() / denom)
        return ops.ce_bwd_sharded(lg, local, lse_global, row_scale).view(ctx.shape), None, None, None, None, None


class MoEExpertsFn(torch.autograd.Function):
    """MixtralExperts.forward (models/mixtral/modeling_mixtral.py:69-93) with its backward, on the routing / gather /
    combine kernels and per-expert GEMMs: pairs (token, k) are sorted by expert, each expert runs its gate|up and down
    projections on its row range, the outputs are un-permuted with the routing weights.

    Backward (all on the same kernels): dY_sorted = w * gather(dOut); dW_down[e] = dY_e^T act_e, dAct = dY_e down[e];
    dGU = glu'(gu) dAct; dW_gu[e] = dGU_e^T x_e, dX_sorted = dGU_e gate_up[e]; dX = combine(dX_sorted, 1);
    d top_k_weights[t,j] = <dOut[t], y_sorted[slot(t,j)]> (the router's gradient path)."""

    @staticmethod
    def forward(ctx, x, top_k_index, top_k_weights, gate_up, down, gelu):
        T, H = x.shape
        k = top_k_index.shape[1]
        E = gate_up.shape[0]
        offsets, slot, tok = ops.moe_route(top_k_index, E)
        xs = ops.moe_gather(x, tok)
        n = xs.shape[0]
        grouped = ops.grouped_ok(gate_up, down)
        off = None
        if grouped:  # one grouped GEMM per projection, expert row ranges read from the device (no host sync in the forward)
            gu = ops.gemm_grouped(xs, gate_up, offsets)
            act = ops.glu_fwd(gu, gelu)
            ys = ops.gemm_grouped(act, down, offsets)
        else:
            off = offsets.tolist()  # per-expert launches: their row ranges are host-side launch parameters
            gu = torch.empty(n, gate_up.shape[1], device=x.device, dtype=x.dtype)
            ys = torch.empty(n, H, device=x.device, dtype=x.dtype)
            for e in range(E):
                lo, hi = off[e], off[e + 1]
                if hi > lo:
                    ops.gemm(xs[lo:hi], gate_up[e], out=gu[lo:hi])
            act = ops.glu_fwd(gu, gelu)
            for e in range(E):
                lo, hi = off[e], off[e + 1]
                if hi > lo:
                    ops.gemm(act[lo:hi], down[e], out=ys[lo:hi])
        out = ops.moe_combine(ys, slot, top_k_weights, T, k)
        ctx.save_for_backward(xs, gu, act, ys, slot, tok, top_k_weights, gate_up, down, offsets)
        ctx.cfg = (off, gelu, T, k, grouped)
        return out

    @staticmethod
    def backward(ctx, dout):
        xs, gu, act, ys, slot, tok, weights, gate_up, down, offsets = ctx.saved_tensors
        off, gelu, T, k, grouped = ctx.cfg
        E = gate_up.shape[0]
        if off is None:  # the weight gradients are per-expert GEMMs over the expert's rows (K-grouped): host-side ranges
            off = offsets.tolist()
        dout = dout.contiguous()
        g = ops.moe_gather(dout, tok)  # dOut of the token behind every sorted slot
        slot_l = slot.long()
        dweights = None
        if ctx.needs_input_grad[2]:
            dots = (g.float() * ys.float()).sum(dim=-1)  # [n] row dot products (n x H elementwise: not a hot op)
            dweights = dots[slot_l].view(T, k).to(weights.dtype)
        w_sorted = torch.empty(slot.numel(), device=g.device, dtype=torch.float32)
        w_sorted[slot_l] = weights.reshape(-1).float()
        dys = (g.float() * w_sorted[:, None]).to(g.dtype)
        dact = ops.gemm_grouped(dys, down, offsets, b_mn=True) if grouped else torch.empty_like(act)
        d_down = torch.zeros_like(down) if ctx.needs_input_grad[4] else None
        for e in range(E):
            lo, hi = off[e], off[e + 1]
            if hi > lo:
                if not grouped:
                    ops.gemm(dys[lo:hi], down[e], b_mn=True, out=dact[lo:hi])
                if d_down is not None:
                    ops.gemm(dys[lo:hi], act[lo:hi], a_mn=True, b_mn=True, out=d_down[e])
        dgu = ops.glu_bwd(dact, gu, gelu)
        dxs = ops.gemm_grouped(dgu, gate_up, offsets, b_mn=True) if grouped else torch.empty_like(xs)
        d_gate_up = torch.zeros_like(gate_up) if ctx.needs_input_grad[3] else None
        for e in range(E):
            lo, hi = off[e], off[e + 1]
            if hi > lo:
                if not grouped:
                    ops.gemm(dgu[lo:hi], gate_up[e], b_mn=True, out=dxs[lo:hi])
                if d_gate_up is not None:
                    ops.gemm(dgu[lo:hi], xs[lo:hi], a_mn=True, b_mn=True, out=d_gate_up[e])
        dx = None
        if ctx.needs_input_grad[0]:
            ones = torch.ones(T, k, device=dxs.device, dtype=torch.float32)
            dx = ops.moe_combine(dxs, slot, ones, T, k)
        return dx, None, dweights, d_gate_up, d_down, None
