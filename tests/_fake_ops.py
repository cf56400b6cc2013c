"""Test-only stand-ins for ``transformers_b200.ops`` so the HOST logic of the kernel path (autograd Functions, fused-weight
handling, tensor-parallel chunking / overlap, shapes and strides handed to the kernels) can run on CPU tensors.

Each function restates the contract documented in ``transformers_b200/ops.py`` with plain torch in the operand dtype
(fp32 in the tests).  ``install()`` swaps them in for the current process only; nothing in the product imports this file,
and the product keeps failing loudly without the CUDA library (tests/test_plugin_cpu.py).
"""
import torch
import torch.nn.functional as F

CALLS = []  # (name, shapes) log so tests can assert what the host logic launched


def _log(name, *ts):
    CALLS.append((name, tuple(tuple(t.shape) for t in ts if t is not None)))


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, accumulate=False):
    assert a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 or a.shape[1] == 1, "gemm A needs a unit inner stride"
    assert b.stride(1) == 1 or b.shape[1] == 1, "gemm B needs a unit inner stride"
    A = a.t() if a_mn else a
    Bm = b if b_mn else b.t()
    r = A @ Bm
    _log("gemm", a, b)
    if out is None:
        assert not accumulate
        return r
    assert out.shape == r.shape and out.stride(1) == 1
    if accumulate:
        out += r
    else:
        out.copy_(r)
    return out


def embedding_fwd(ids, weight, scale=None):
    y = F.embedding(ids, weight)
    _log("embedding_fwd", ids)
    return y * torch.tensor(scale, dtype=weight.dtype) if scale is not None else y


def embedding_bwd(ids, dout, num_embeddings, padding_idx, scale=None):
    H = dout.shape[-1]
    d = dout.reshape(-1, H)
    if scale is not None:
        d = d * torch.tensor(scale, dtype=d.dtype)
    dw = torch.zeros(num_embeddings, H, dtype=dout.dtype)
    dw.index_add_(0, ids.reshape(-1), d)
    if padding_idx is not None:
        dw[padding_idx] = 0
    _log("embedding_bwd", ids)
    return dw


def rmsnorm_fwd(x, weight, eps, gemma=False, residual=None):
    r = x if residual is None else x + residual
    rf = r.float()
    rstd = torch.rsqrt(rf.pow(2).mean(-1) + eps)
    wv = (1.0 + weight.float()) if gemma else weight.float()
    y = (rf * rstd[..., None] * wv).to(x.dtype)
    _log("rmsnorm_fwd", x)
    return y, rstd.reshape(-1), (r if residual is not None else None)


def rmsnorm_bwd(dy, x, weight, rstd, gemma=False):
    H = x.shape[-1]
    xf, dyf = x.reshape(-1, H).float(), dy.reshape(-1, H).float()
    wv = (1.0 + weight.float()) if gemma else weight.float()
    xhat = xf * rstd[:, None]
    g = dyf * wv
    dx = rstd[:, None] * (g - xhat * (g * xhat).mean(-1, keepdim=True))
    dw = (dyf * xhat).sum(0)
    _log("rmsnorm_bwd", x)
    return dx.to(x.dtype).view(x.shape), dw.to(weight.dtype)


def rope_table(inv_freq, position_ids, attention_scaling=1.0, dtype=torch.bfloat16):
    ang = position_ids.to(torch.float32)[:, :, None] * inv_freq.to(torch.float32)[None, None, :]
    emb = torch.cat([ang, ang], dim=-1)
    _log("rope_table", position_ids)
    return (emb.cos() * attention_scaling).to(dtype), (emb.sin() * attention_scaling).to(dtype)


def rope_(qkv, cos, sin, n_rot_heads, head_dim, backward=False):
    B, S, W = qkv.shape
    assert qkv.is_contiguous()
    D, h = head_dim, head_dim // 2
    x = qkv[..., : n_rot_heads * D].reshape(B, S, n_rot_heads, D)
    c = cos.reshape(-1, S, 1, D)
    s = sin.reshape(-1, S, 1, D)
    if not backward:
        rot = torch.cat([-x[..., h:], x[..., :h]], -1)
        y = x * c + rot * s
    else:
        z = x * s
        y = x * c + torch.cat([z[..., h:], -z[..., :h]], -1)
    qkv[..., : n_rot_heads * D] = y.reshape(B, S, n_rot_heads * D)
    _log("rope_", qkv)
    return qkv


def add(a, b):
    _log("add", a)
    return a + b


def _act(g, gelu):
    return F.gelu(g, approximate="tanh") if gelu else F.silu(g)


GLU_BLOCK = 128


def _halves(gu, interleaved):
    """(gate, up) views of a [.., 2I] projection output in the plain or the 128-column block-interleaved layout."""
    I = gu.shape[-1] // 2
    if not interleaved:
        return gu[..., :I], gu[..., I:]
    v = gu.reshape(*gu.shape[:-1], I // GLU_BLOCK, 2, GLU_BLOCK)
    return v[..., 0, :].reshape(*gu.shape[:-1], I), v[..., 1, :].reshape(*gu.shape[:-1], I)


def glu_fwd(gu, gelu=False, interleaved=False):
    g, u = _halves(gu, interleaved)
    _log("glu_fwd", gu)
    return _act(g, gelu) * u


def glu_bwd(dh, gu, gelu=False, interleaved=False):
    with torch.enable_grad():
        x = gu.detach().clone().requires_grad_(True)
        g, u = _halves(x, interleaved)
        y = _act(g, gelu) * u
        (dgu,) = torch.autograd.grad(y, x, dh.reshape(y.shape))
    _log("glu_bwd", gu)
    return dgu


def interleave_gate_up(wg, wu):
    I, K = wg.shape
    assert I % GLU_BLOCK == 0
    return torch.stack([wg.reshape(I // GLU_BLOCK, GLU_BLOCK, K), wu.reshape(I // GLU_BLOCK, GLU_BLOCK, K)], dim=1).reshape(2 * I, K).contiguous()


def deinterleave_gate_up(t):
    n = t.shape[0] // (2 * GLU_BLOCK)
    v = t.reshape(n, 2, GLU_BLOCK, *t.shape[1:])
    return v[:, 0].reshape(n * GLU_BLOCK, *t.shape[1:]), v[:, 1].reshape(n * GLU_BLOCK, *t.shape[1:])


def gemm_glu(a, w_ilv, gelu=False, gu_out=None, h_out=None):
    gu = a @ w_ilv.t()
    g, u = _halves(gu, True)
    h = _act(g, gelu) * u
    _log("gemm_glu", a, w_ilv)
    if gu_out is not None:
        gu_out.copy_(gu)
        gu = gu_out
    if h_out is not None:
        h_out.copy_(h)
        h = h_out
    return gu, h


def glu_fusable(M, I):
    return M > 128 and I % GLU_BLOCK == 0


def gemm_grouped(a, b, offsets, *, b_mn=False, out=None):
    assert offsets.dtype == torch.int32 and offsets.numel() == b.shape[0] + 1
    N = b.shape[2] if b_mn else b.shape[1]
    if out is None:
        out = a.new_zeros(a.shape[0], N)
    off = offsets.tolist()
    for g in range(b.shape[0]):
        lo, hi = off[g], off[g + 1]
        if hi > lo:
            out[lo:hi] = a[lo:hi] @ (b[g] if b_mn else b[g].t())
    _log("gemm_grouped", a, b)
    return out


def grouped_ok(gate_up, down):
    E, I2, H = gate_up.shape
    return E <= 64 and I2 % 64 == 0 and H % 64 == 0 and down.shape[2] % 64 == 0


def _attn_math(q, k, v, scale, causal, window, softcap, kv_start, kv_end):
    B, Sq, Hq, D = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    rep = Hq // Hkv
    qf = q.permute(0, 2, 1, 3).double()
    kf = k.permute(0, 2, 1, 3).repeat_interleave(rep, 1).double()
    vf = v.permute(0, 2, 1, 3).repeat_interleave(rep, 1).double()
    s = qf @ kf.transpose(-1, -2) * scale
    if softcap:
        s = softcap * torch.tanh(s / softcap)
    qpos = torch.arange(Sq) + (Skv - Sq)  # bottom-right aligned
    kidx = torch.arange(Skv)
    m = torch.ones(Sq, Skv, dtype=torch.bool)
    if causal:
        m &= kidx[None, :] <= qpos[:, None]
    if window:
        m &= kidx[None, :] > qpos[:, None] - window
    m = m[None, None].expand(B, 1, Sq, Skv)
    if kv_start is not None:
        m = m & (kidx[None, None, None, :] >= kv_start.long()[:, None, None, None])
    if kv_end is not None:
        m = m & (kidx[None, None, None, :] < kv_end.long()[:, None, None, None])
    s = s.masked_fill(~m, float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.exp(s - lse[..., None])
    p = torch.nan_to_num(p, nan=0.0)  # fully masked rows -> zeros
    out = (p @ vf).permute(0, 2, 1, 3)
    return out, lse


def attn_fwd(q, k, v, *, scale, causal, window=0, softcap=0.0, kv_start=None, kv_end=None, out=None):
    for t in (q, k, v):
        assert t.dim() == 4 and t.stride(3) == 1, "attention operands: [B,S,h,D] views with unit stride on D"
    o, lse = _attn_math(q, k, v, scale, causal, window, softcap, kv_start, kv_end)
    B, Sq, Hq, _ = q.shape
    ls = (Sq + 127) // 128 * 128
    lse_p = torch.zeros(B, Hq, ls, dtype=torch.float32)
    lse_p[..., :Sq] = lse.float()
    o = o.to(q.dtype).contiguous()
    if out is not None:
        out.copy_(o)
        o = out
    _log("attn_fwd", q, k, v)
    return o, lse_p


def attn_bwd(q, k, v, out, dout, lse, dq, dk, dv, *, scale, causal, window=0, softcap=0.0, kv_start=None, kv_end=None):
    for t in (q, k, v, out, dout, dq, dk, dv):
        assert t.stride(3) == 1
    with torch.enable_grad():
        qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        o, _ = _attn_math(qq, kk, vv, scale, causal, window, softcap, kv_start, kv_end)
        gq, gk, gv = torch.autograd.grad(o, (qq, kk, vv), dout.double())
    dq.copy_(gq.to(dq.dtype))
    dk.copy_(gk.to(dk.dtype))
    dv.copy_(gv.to(dv.dtype))
    _log("attn_bwd", q, k, v)
    return dq, dk, dv


def _ce_targets(labels, B, S, shift, ignore_index):
    labels = labels.reshape(B, S)
    if shift:
        tgt = torch.full_like(labels, ignore_index)
        tgt[:, :-1] = labels[:, 1:]
    else:
        tgt = labels
    return tgt.reshape(-1)


def ce_fwd(logits, labels, shift=True, ignore_index=-100, num_items=None):
    B, S, V = logits.shape
    tgt = _ce_targets(labels, B, S, shift, ignore_index)
    lg = logits.reshape(B * S, V).float()
    lse = torch.logsumexp(lg, -1)
    valid = tgt != ignore_index
    nll = (lse - lg.gather(1, tgt.clamp(min=0)[:, None])[:, 0]) * valid
    denom = torch.tensor([float(num_items) if num_items else float(valid.sum())])
    _log("ce_fwd", logits)
    return (nll.sum() / denom[0]).to(torch.float32), lse, denom


def ce_bwd(logits, labels, lse, dloss, denom, shift=True, ignore_index=-100, out=None):
    B, S, V = logits.shape
    tgt = _ce_targets(labels, B, S, shift, ignore_index)
    lg = logits.reshape(B * S, V).float()
    p = torch.exp(lg - lse[:, None])
    valid = tgt != ignore_index
    p[torch.arange(B * S)[valid], tgt[valid]] -= 1.0
    p = p * valid[:, None] * (dloss.reshape(()).float() / denom[0])
    _log("ce_bwd", logits)
    if out is not None:
        out.copy_(p.to(logits.dtype))
        return out.view(B, S, V)
    return p.to(logits.dtype).view(B, S, V)


def ce_row_lse(logits):
    _log("ce_row_lse", logits)
    return torch.logsumexp(logits.float(), -1)


def ce_bwd_sharded(logits, target_local, lse_global, row_scale):
    T, V = logits.shape
    p = torch.exp(logits.float() - lse_global[:, None])
    mine = (target_local >= 0) & (target_local < V)
    p[torch.arange(T)[mine], target_local[mine]] -= 1.0
    _log("ce_bwd_sharded", logits)
    return (p * row_scale[:, None]).to(logits.dtype)


def moe_route(top_k_index, E):
    T, k = top_k_index.shape
    flat = top_k_index.reshape(-1)
    order = torch.sort(flat, stable=True).indices  # any order inside an expert is allowed; stable = deterministic
    n = flat.numel()
    slot = torch.empty(n, dtype=torch.int32)
    slot[order] = torch.arange(n, dtype=torch.int32)
    tok = (order // k).to(torch.int32)
    counts = torch.bincount(flat, minlength=E)
    offsets = torch.zeros(E + 1, dtype=torch.int32)
    offsets[1:] = counts.cumsum(0).to(torch.int32)
    _log("moe_route", top_k_index)
    return offsets, slot, tok


def moe_gather(x, tok):
    _log("moe_gather", x)
    return x[tok.long()].contiguous()


def moe_combine(ys, slot, weights, T, k):
    out = (ys[slot.long()].view(T, k, -1).float() * weights.float()[..., None]).sum(1)
    _log("moe_combine", ys)
    return out.to(ys.dtype)


def moe_experts_forward(x, top_k_index, top_k_weights, gate_up, down, gelu=False):
    E = gate_up.shape[0]
    offsets, slot, tok = moe_route(top_k_index, E)
    off = offsets.tolist()
    xs = moe_gather(x, tok)
    ys = torch.empty(xs.shape[0], x.shape[1], dtype=x.dtype)
    for e in range(E):
        lo, hi = off[e], off[e + 1]
        if hi > lo:
            ys[lo:hi] = glu_fwd(xs[lo:hi] @ gate_up[e].t(), gelu) @ down[e].t()
    return moe_combine(ys, slot, top_k_weights, *top_k_index.shape)


def kv_append(k_new, v_new, k_cache, v_cache, offset):
    B, H, q, D = k_new.shape
    assert k_new.stride(3) == 1 and v_new.stride(3) == 1 and k_cache.stride() == v_cache.stride()
    assert offset + q <= k_cache.shape[2], "append beyond the cache capacity"
    if k_new.untyped_storage().data_ptr() == k_cache.untyped_storage().data_ptr():  # in-place move inside one buffer
        src0 = (k_new.data_ptr() - k_cache.data_ptr()) // (k_cache.element_size() * k_cache.stride(2))
        assert src0 >= offset + q or src0 + q <= offset, "kv_append: source and destination rows overlap (parallel copy)"
    k_cache[:, :, offset:offset + q] = k_new
    v_cache[:, :, offset:offset + q] = v_new
    _log("kv_append", k_new)


# ---- optimizer: the "device pointers" of the tables are resolved through a registry of live CPU tensors
_PTRS = {}


def register_tensors(*ts):
    for t in ts:
        _PTRS[t.data_ptr()] = t


def optim_chunk_elems():
    return 32768


def _rows(table):
    return [tuple(int(x) for x in r) for r in table.tolist()]


def adamw_step(table, chunk_map, *, state_fp32, master=False, lr, beta1, beta2, eps, weight_decay, bias_correction1,
               bias_correction2_sqrt, grad_scale=None):
    chunk = optim_chunk_elems()
    seen = {}
    for ti, ci in chunk_map.tolist():
        seen.setdefault(ti, []).append(ci)
    gs = float(grad_scale[0]) if grad_scale is not None else 1.0
    for ti, (pp, gp, mp, vp, n, mw) in enumerate(_rows(table)):
        assert seen.get(ti) == list(range((n + chunk - 1) // chunk)), "chunk map must cover every tensor exactly once"
        p, g, m, v = _PTRS[pp], _PTRS[gp], _PTRS[mp], _PTRS[vp]
        assert p.numel() == n and (m.dtype == torch.float32) == bool(state_fp32)
        assert bool(mw) == bool(master)
        pf, gf, mf, vf = (_PTRS[mw].float() if master else p.float()), g.float() * gs, m.float(), v.float()
        pf = pf - lr * weight_decay * pf
        mf = mf + (gf - mf) * (1.0 - beta1)
        vf = beta2 * vf + (1.0 - beta2) * gf * gf
        pf = pf - (lr / bias_correction1) * (mf / (vf.sqrt() / bias_correction2_sqrt + eps))
        if master:
            _PTRS[mw].copy_(pf)
        p.copy_(pf)
        m.copy_(mf)
        v.copy_(vf)
    _log("adamw_step", table)


def grad_norm(table, chunk_map, max_norm=0.0):
    tot = sum(float((_PTRS[gp].double() ** 2).sum()) for _, gp, _, _, _, _ in _rows(table)) ** 0.5
    coef = min(1.0, max_norm / (tot + 1e-6)) if max_norm > 0 else 1.0
    _log("grad_norm", table)
    return torch.tensor([tot, coef], dtype=torch.float32)


def grad_scale_(table, chunk_map, coef):
    for _, gp, _, _, _, _ in _rows(table):
        _PTRS[gp].mul_(float(coef[0]))
    _log("grad_scale_", table)


def pull_reduce(peer_ptrs, offset_elems, n_elems, out, residual=None):
    acc = torch.zeros(n_elems, dtype=torch.float32)
    if residual is not None:
        acc += residual.reshape(-1).float()
    for p in peer_ptrs:  # rank order, like the kernel
        acc += _PTRS[p].reshape(-1)[offset_elems:offset_elems + n_elems].float()
    out.copy_(acc.view(out.shape))
    _log("pull_reduce", out)
    return out


def gemm_scatter(a, b, dest_ptrs, rank, *, a_mn=False, b_mn=False):
    r = (a.t() if a_mn else a) @ (b if b_mn else b.t())
    rows = r.shape[0] // len(dest_ptrs)
    assert rows % 256 == 0, "scatter epilogue: whole 256-row tiles per owner"
    for o, key in enumerate(dest_ptrs):
        _PTRS[key].copy_(r[o * rows:(o + 1) * rows])
    _log("gemm_scatter", a, b)


class FakePeerWorkspace:
    """Stand-in for transformers_b200.symm.PeerWorkspace over gloo: "peer-mapped" buffers are emulated by all-gathering
    every rank's buffer at the barrier.  Checks the double-buffer protocol the real workspace relies on."""

    def __init__(self, group=None, dtype=torch.float32, scatter_epilogue=False):
        import contextlib

        self.scatter_epilogue = scatter_epilogue
        self._outbox = None

        import torch.distributed as dist

        self.dist, self.group, self.dtype = dist, group, dtype
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._partial, self._peers_partial, self._peers_shard = None, None, None
        self._null = contextlib.nullcontext
        self.ops = 0

    def next_partial(self, rows, cols):
        self._partial = torch.empty(rows, cols, dtype=self.dtype)
        self._peers_partial = None
        return self._partial

    def _gather(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t.contiguous(), group=self.group)
        return parts

    def next_staging(self, rows, cols):
        # outbox[r] = what this rank's GEMM epilogue writes into rank r's slot; delivered at the barrier below
        self._outbox = torch.zeros(self.world, rows, cols, dtype=self.dtype)
        self._slots = None
        dest = []
        for r in range(self.world):
            _PTRS[self._outbox[r].data_ptr()] = self._outbox[r]
            dest.append(self._outbox[r].data_ptr())
        self._slot_keys = [("slot", id(self), s) for s in range(self.world)]
        return dest, self._slot_keys

    def publish_partial(self):
        if self._outbox is not None:  # scatter epilogue: emulate the NVLink stores with an all-to-all of the outboxes
            boxes = self._gather(self._outbox)  # boxes[src][dst]
            for s in range(self.world):
                _PTRS[self._slot_keys[s]] = boxes[s][self.rank].contiguous()
            self._outbox = None
        else:
            self._peers_partial = self._gather(self._partial)
        self.ops += 1

    def partial_ptrs(self):
        assert self._peers_partial is not None, "pull before publish_partial(): the barrier is what makes peer data visible"
        keys = []
        for t in self._peers_partial:
            _PTRS[t.data_ptr()] = t
            keys.append(t.data_ptr())
        return keys

    def publish_shard(self, local2):
        self._peers_shard = self._gather(local2)
        self.ops += 1

    def peer_shard(self, src, rows, cols):
        assert src != self.rank, "own rows are a local copy"
        return self._peers_shard[src].view(rows, cols)

    def copy_context(self, i=0):
        return self._null()

    def join_copies(self):
        pass


_NAMES = ["gemm", "gemm_grouped", "grouped_ok", "gemm_glu", "glu_fusable", "interleave_gate_up", "deinterleave_gate_up", "embedding_fwd", "embedding_bwd", "rmsnorm_fwd", "rmsnorm_bwd", "rope_", "rope_table", "glu_fwd", "glu_bwd", "attn_fwd",
          "attn_bwd", "ce_fwd", "ce_bwd", "ce_row_lse", "ce_bwd_sharded", "add", "kv_append", "pull_reduce", "gemm_scatter", "moe_route", "moe_gather",
          "moe_combine", "moe_experts_forward", "optim_chunk_elems",
          "adamw_step", "grad_norm", "grad_scale_"]


def install(setattr_fn=setattr):
    """Swap the fakes in (process-local).  ``setattr_fn`` may be pytest's ``monkeypatch.setattr`` so the swap is undone."""
    import transformers_b200.modules as M
    import transformers_b200.ops as ops

    g = globals()
    for n in _NAMES:
        setattr_fn(ops, n, g[n])
    setattr_fn(M, "_on_b200", lambda t: True)
    setattr_fn(M, "KERNEL_DTYPES", (torch.bfloat16, torch.float32))
    CALLS.clear()
