"""The ALGORITHM of csrc/attention_decode.cu restated in torch on CPU (same split ranges, same per-(warp, row-group) online
softmax in the exp2 domain, same two-level merge through (m, l, O) triples) against a plain fp32 softmax.  This cannot
check the CUDA mechanics (that is tests/test_kernels2_gpu.py on a device); it pins the index arithmetic and the merge
formulas the kernel implements, including empty splits, padding ranges, the sliding window and the softcap."""
import math

import pytest
import torch

LOG2E = 1.4426950408889634
DEC_WARPS = 4


def _splits(B, Hkv, Skv, sms=148):  # b200_attn_decode_splits
    ctas = max(B * Hkv, 1)
    n = (2 * sms + ctas - 1) // ctas
    return max(1, min(n, (Skv + 255) // 256, 64))


def decode_like_the_kernel(q, k, v, scale, softcap, window, kv_start, kv_end, nsplit):
    B, Hq, D = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    G = Hq // Hkv
    rpi = 32 // (D // 8)
    out = torch.zeros(B, Hq, D)
    lse = torch.full((B, Hq), float("-inf"))
    pre = scale / softcap if softcap > 0 else scale * LOG2E
    for b in range(B):
        lo = int(kv_start[b]) if kv_start is not None else 0
        hi = min(int(kv_end[b]) if kv_end is not None else Skv, Skv)
        if window > 0 and lo < Skv - window:
            lo = Skv - window
        lo = max(lo, 0)
        span = max(hi - lo, 0)
        per = (span + nsplit - 1) // nsplit
        for hkv in range(Hkv):
            ws = []  # (m [G], l [G], O [G, D]) per split
            qs = q[b, hkv * G:(hkv + 1) * G].float() * pre
            for split in range(nsplit):
                r0 = lo + split * per
                r1 = min(hi, r0 + per)
                states = []
                for warp in range(DEC_WARPS):
                    for sub in range(rpi):
                        m = torch.full((G,), float("-inf"))
                        l = torch.zeros(G)
                        acc = torch.zeros(G, D)
                        rb = r0 + warp * rpi
                        while rb < r1:
                            r = rb + sub
                            if r < r1:
                                s = qs @ k[b, r, hkv].float()
                                if softcap > 0:
                                    s = softcap * LOG2E * torch.tanh(s)
                                mn = torch.maximum(m, s)
                                corr = torch.exp2(m - mn)
                                pr = torch.exp2(s - mn)
                                l = l * corr + pr
                                acc = acc * corr[:, None] + pr[:, None] * v[b, r, hkv].float()[None]
                                m = mn
                            rb += DEC_WARPS * rpi
                        states.append((m, l, acc))
                ms = torch.stack([s_[0] for s_ in states])
                mm = ms.max(0).values
                w = torch.where(torch.isinf(ms), torch.zeros_like(ms), torch.exp2(ms - mm))
                ws.append((mm, (torch.stack([s_[1] for s_ in states]) * w).sum(0),
                           (torch.stack([s_[2] for s_ in states]) * w[..., None]).sum(0)))
            ms = torch.stack([t[0] for t in ws])
            mm = ms.max(0).values
            w = torch.where(torch.isinf(ms), torch.zeros_like(ms), torch.exp2(ms - mm))
            ll = (torch.stack([t[1] for t in ws]) * w).sum(0)
            oo = (torch.stack([t[2] for t in ws]) * w[..., None]).sum(0)
            inv = torch.where(ll > 0, 1.0 / ll, torch.zeros_like(ll))
            out[b, hkv * G:(hkv + 1) * G] = oo * inv[:, None]
            lse[b, hkv * G:(hkv + 1) * G] = torch.where(ll > 0, mm * math.log(2.0) + torch.log(ll), torch.full_like(ll, float("-inf")))
    return out, lse


@pytest.mark.parametrize("B,Hq,Hkv,D,ctx,window,softcap,nsplit", [
    (2, 4, 2, 64, 37, 0, 0.0, 3), (1, 8, 1, 128, 70, 16, 0.0, 5), (2, 2, 2, 256, 9, 0, 30.0, 4), (1, 4, 4, 64, 1, 0, 0.0, 2),
    (2, 4, 2, 128, 40, 0, 0.0, None),
])
def test_decode_algorithm_matches_softmax(B, Hq, Hkv, D, ctx, window, softcap, nsplit):
    torch.manual_seed(ctx + D)
    q = torch.randn(B, Hq, D)
    k = torch.randn(B, ctx, Hkv, D)
    v = torch.randn(B, ctx, Hkv, D)
    kv_start = torch.tensor([0] + [3] * (B - 1)) if ctx > 8 else None
    kv_end = torch.tensor([ctx] + [ctx - 2] * (B - 1)) if ctx > 8 else None
    scale = D ** -0.5
    n = nsplit if nsplit is not None else _splits(B, Hkv, ctx)
    out, lse = decode_like_the_kernel(q, k, v, scale, softcap, window, kv_start, kv_end, n)
    G = Hq // Hkv
    s = torch.einsum("bhd,bkhd->bhk", q, k.repeat_interleave(G, dim=2)) * scale
    if softcap:
        s = softcap * torch.tanh(s / softcap)
    idx = torch.arange(ctx)
    valid = torch.ones(B, ctx, dtype=torch.bool)
    if window:
        valid &= idx[None] >= ctx - window
    if kv_start is not None:
        valid &= (idx[None] >= kv_start[:, None]) & (idx[None] < kv_end[:, None])
    s = s.masked_fill(~valid[:, None], float("-inf"))
    want = torch.einsum("bhk,bkhd->bhd", torch.softmax(s, -1), v.repeat_interleave(G, dim=2))
    torch.testing.assert_close(out, want, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(lse, torch.logsumexp(s, -1), atol=1e-5, rtol=1e-5)


def test_fully_masked_batch_row_gives_zeros():
    q, k, v = torch.randn(1, 2, 64), torch.randn(1, 12, 2, 64), torch.randn(1, 12, 2, 64)
    out, lse = decode_like_the_kernel(q, k, v, 0.125, 0.0, 0, torch.tensor([12]), torch.tensor([12]), 3)
    assert torch.count_nonzero(out) == 0 and torch.isinf(lse).all()
