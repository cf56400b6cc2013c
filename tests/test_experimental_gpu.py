"""GPU checks of kernels that were written after the round's GPU budget was spent (never run on a device yet).  Opt-in so
an unvalidated kernel cannot turn the regular `-m gpu` suite red: B200_EXPERIMENTAL=1 python -m pytest tests -m gpu."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.environ.get("B200_EXPERIMENTAL"), reason="set B200_EXPERIMENTAL=1")]


def test_ce_sharded_matches_full_ce():
    """Two vocabulary shards combined by hand must reproduce b200_ce_fwd / b200_ce_bwd on the full logits bit for bit in the
    statistics (same fp32 reductions per shard) and to bf16 rounding in the gradient."""
    from transformers_b200 import ops

    torch.manual_seed(0)
    B, S, V, N = 2, 64, 4096, 2
    logits = (torch.randn(B, S, V, device="cuda") * 3).to(torch.bfloat16)
    labels = torch.randint(0, V, (B, S), device="cuda")
    labels[0, :5] = -100
    loss, lse, denom = ops.ce_fwd(logits, labels, shift=True)
    dl = ops.ce_bwd(logits, labels, lse, torch.ones((), device="cuda"), denom, shift=True)
    tgt = torch.full_like(labels, -100)
    tgt[:, :-1] = labels[:, 1:]
    tgt = tgt.reshape(-1)
    valid = tgt != -100
    shards = logits.reshape(B * S, V).chunk(N, dim=-1)
    lses = torch.stack([ops.ce_row_lse(s.contiguous()) for s in shards])
    lse_g = torch.logsumexp(lses, 0)
    torch.testing.assert_close(lse_g, lse, atol=1e-5, rtol=1e-6)
    scale = valid.float() / denom
    for r, s in enumerate(shards):
        local = tgt - r * (V // N)
        local = torch.where(valid & (local >= 0) & (local < V // N), local, torch.full_like(local, -1))
        d = ops.ce_bwd_sharded(s.contiguous(), local, lse, scale)
        torch.testing.assert_close(d.float(), dl.reshape(B * S, V).chunk(N, dim=-1)[r].float(), atol=1e-6, rtol=1e-2)
