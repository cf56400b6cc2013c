"""Optimizer step after the hot path on CPU:
  * the oracle (oracle/adamw_oracle.py) is pinned against torch.optim.AdamW / torch.nn.utils.clip_grad_norm_ themselves --
    the code the reference's Trainer calls (trainer.py:1788, :2538-2542; trainer_optimizer.py:201-208);
  * the host logic of transformers_b200.optim (param groups, per-step bias corrections, multi-tensor tables and chunk maps,
    fused clipping, state layout) runs against torch.optim.AdamW with the kernels replaced by tests/_fake_ops.py."""
import math

import pytest
import torch

import _fake_ops
from oracle import adamw_oracle as O

SHAPES = [(7,), (64, 33), (3, 5, 8), (40000,)]  # odd sizes, one tensor larger than a 32768-element chunk


def _params(dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(s, generator=g) * 0.5).to(dtype)) for s in SHAPES]


def _set_grads(ps, step, scale=1.0):
    g = torch.Generator().manual_seed(100 + step)
    for p in ps:
        p.grad = (torch.randn(p.shape, generator=g) * scale).to(p.dtype)


@pytest.mark.parametrize("wd", [0.0, 0.1])
def test_oracle_matches_torch_adamw_fp32(wd):
    ps = _params(torch.float32)
    opt = torch.optim.AdamW(ps, lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd, foreach=False, fused=False)
    mine = [(p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p)) for p in ps]
    for step in range(1, 8):
        _set_grads(ps, step)
        opt.step()
        mine = [O.adamw_step(p, q.grad, m, v, step, 1e-2, 0.9, 0.95, 1e-8, wd) for (p, m, v), q in zip(mine, ps)]
    for (p, m, v), q in zip(mine, ps):
        torch.testing.assert_close(p, q.detach(), atol=2e-6, rtol=2e-6)
        torch.testing.assert_close(m, opt.state[q]["exp_avg"], atol=1e-7, rtol=1e-6)
        torch.testing.assert_close(v, opt.state[q]["exp_avg_sq"], atol=1e-7, rtol=1e-6)


def test_oracle_tracks_torch_adamw_bf16_within_rounding():
    ps = _params(torch.bfloat16)
    opt = torch.optim.AdamW(ps, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, foreach=False, fused=False)
    mine = [(p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p)) for p in ps]
    for step in range(1, 6):
        _set_grads(ps, step)
        opt.step()
        mine = [O.adamw_step(p, q.grad, m, v, step, 1e-2, 0.9, 0.999, 1e-8, 0.01) for (p, m, v), q in zip(mine, ps)]
    for (p, m, v), q in zip(mine, ps):
        # bf16 has 8 bits of mantissa: per-op rounding (torch CPU) vs one rounding per stored tensor (oracle / kernel)
        torch.testing.assert_close(p.float(), q.detach().float(), atol=4e-3, rtol=2e-2)
        torch.testing.assert_close(m.float(), opt.state[q]["exp_avg"].float(), atol=4e-3, rtol=3e-2)


def test_oracle_clip_matches_torch():
    ps = _params(torch.float32)
    _set_grads(ps, 1, scale=3.0)
    total, coef = O.grad_norm_and_coef([p.grad for p in ps], 1.0)
    ref = torch.nn.utils.clip_grad_norm_(ps, 1.0)
    assert abs(total - float(ref)) < 1e-3 * total and coef < 1.0
    after = math.sqrt(sum(float((p.grad.double() ** 2).sum()) for p in ps))
    assert abs(after - total * coef) < 1e-4


# ------------------------------------------------------------------------------------------------ host logic of B200AdamW
@pytest.fixture
def fakes(monkeypatch):
    import transformers_b200.optim as optim

    _fake_ops.install(monkeypatch.setattr)
    monkeypatch.setattr(optim, "_DTYPES", (torch.bfloat16, torch.float32))
    return optim


def _register(opt):
    for group in opt.param_groups:
        for p in group["params"]:
            if p.grad is not None:
                st = opt._init_state(p)
                _fake_ops.register_tensors(p, p.grad, st["exp_avg"], st["exp_avg_sq"])


@pytest.mark.parametrize("max_norm", [None, 0.5])
def test_b200_adamw_host_logic_matches_torch(fakes, max_norm):
    a, b = _params(torch.float32), _params(torch.float32)
    groups = lambda ps: [{"params": ps[:2], "weight_decay": 0.1}, {"params": ps[2:], "weight_decay": 0.0}]  # Trainer's decay / no-decay split
    ref = torch.optim.AdamW(groups(a), lr=3e-3, betas=(0.9, 0.95), eps=1e-8)
    ours = fakes.B200AdamW(groups(b), lr=3e-3, betas=(0.9, 0.95), eps=1e-8, max_grad_norm=max_norm, fused=True)
    sched_a = torch.optim.lr_scheduler.LambdaLR(ref, lambda s: 1.0 / (1 + s))
    sched_b = torch.optim.lr_scheduler.LambdaLR(ours, lambda s: 1.0 / (1 + s))
    for step in range(1, 6):
        _set_grads(a, step, scale=2.0)
        _set_grads(b, step, scale=2.0)
        if step == 3:  # a parameter without a gradient is skipped, like torch
            a[1].grad = None
            b[1].grad = None
        if max_norm:
            n_ref = torch.nn.utils.clip_grad_norm_(a, max_norm)
        _register(ours)
        _fake_ops.CALLS.clear()
        ref.step()
        ours.step()
        sched_a.step()
        sched_b.step()
        names = [c[0] for c in _fake_ops.CALLS]
        assert names.count("adamw_step") == (3 if step >= 4 else 2)  # one launch per group (+1 once p[1] lags a step behind)
        if max_norm:
            assert names.count("grad_norm") == 1 and "grad_scale_" not in names  # clip fused into the update
            torch.testing.assert_close(ours.grad_norm, n_ref.float(), rtol=1e-5, atol=1e-6)
    for p, q in zip(a, b):
        torch.testing.assert_close(q.detach(), p.detach(), atol=3e-6, rtol=1e-5)
        assert set(ours.state[q]) == set(ref.state[p]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(ours.state[q]["step"]) == float(ref.state[p]["step"])
        torch.testing.assert_close(ours.state[q]["exp_avg_sq"], ref.state[p]["exp_avg_sq"], atol=1e-6, rtol=1e-5)
    sd = ours.state_dict()
    assert sd["param_groups"][0]["weight_decay"] == 0.1 and len(sd["state"]) == len(SHAPES)


def test_clip_grad_norm_host_logic_matches_torch(fakes):
    a, b = _params(torch.float32), _params(torch.float32)
    _set_grads(a, 1, scale=3.0)
    _set_grads(b, 1, scale=3.0)
    b[2].grad = None
    a[2].grad = None
    _fake_ops.register_tensors(*[p.grad for p in b if p.grad is not None])
    n_ref = torch.nn.utils.clip_grad_norm_(a, 1.0)
    n = fakes.clip_grad_norm_(b, 1.0)
    torch.testing.assert_close(n, n_ref.float(), rtol=1e-5, atol=1e-6)
    for p, q in zip(a, b):
        if p.grad is not None:
            torch.testing.assert_close(q.grad, p.grad, rtol=1e-5, atol=1e-7)


def test_b200_adamw_rejects_what_it_cannot_do(fakes):
    from transformers_b200 import B200Error

    ps = _params(torch.float32)
    with pytest.raises(B200Error):
        fakes.B200AdamW(ps, amsgrad=True)
    with pytest.raises(TypeError):
        fakes.B200AdamW(ps, nesterov=True)
    opt = fakes.B200AdamW([torch.nn.Parameter(torch.zeros(4, dtype=torch.float16))])
    opt.param_groups[0]["params"][0].grad = torch.zeros(4, dtype=torch.float16)
    with pytest.raises(B200Error):
        opt.step()


def test_optimizer_fails_loudly_without_a_device():
    """No CPU fallback: without the fakes the real ops refuse to run off a B200."""
    from transformers_b200 import B200Error
    from transformers_b200.optim import B200AdamW

    p = torch.nn.Parameter(torch.zeros(16, dtype=torch.bfloat16))
    p.grad = torch.ones(16, dtype=torch.bfloat16)
    with pytest.raises(B200Error):
        B200AdamW([p]).step()


def test_b200_adamw_master_weights_follow_fp32_adamw(fakes):
    """bf16 parameters + fp32 masters: the masters must track torch.optim.AdamW run in fp32 on the same gradients, while a
    plain bf16 run loses the small updates."""
    a = _params(torch.float32)
    b = [torch.nn.Parameter(p.detach().to(torch.bfloat16)) for p in a]
    for p, q in zip(a, b):
        p.data.copy_(q.detach().float())  # same starting point (bf16-representable)
    ref = torch.optim.AdamW(a, lr=1e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    ours = fakes.B200AdamW(b, lr=1e-5, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0, master_weights=True, state_dtype=torch.float32)
    for step in range(1, 9):
        _set_grads(b, step)
        for p, q in zip(a, b):
            p.grad = q.grad.float()
        for group in ours.param_groups:
            for q in group["params"]:
                st = ours._init_state(q)
                _fake_ops.register_tensors(q, q.grad, st["exp_avg"], st["exp_avg_sq"], st["master"])
        ref.step()
        ours.step()
    for p, q in zip(a, b):
        master = ours.state[q]["master"]
        torch.testing.assert_close(master, p.detach(), atol=1e-6, rtol=1e-5)
        assert torch.equal(q.detach(), master.to(torch.bfloat16))
    assert "master" in ours.state_dict()["state"][0]


def test_state_dict_round_trip_keeps_state_dtypes(fakes):
    """torch's Optimizer.load_state_dict casts floating-point state to the parameter dtype (bf16); the fp32 master copy and
    fp32 moments must come back as fp32 (the kernel reads them as float*), bit for bit."""
    b = [torch.nn.Parameter(p.detach().to(torch.bfloat16)) for p in _params(torch.float32)]
    ours = fakes.B200AdamW(b, lr=1e-3, master_weights=True, state_dtype=torch.float32)
    _set_grads(b, 1)
    for q in b:
        st = ours._init_state(q)
        _fake_ops.register_tensors(q, q.grad, st["exp_avg"], st["exp_avg_sq"], st["master"])
    ours.step()
    sd = ours.state_dict()
    fresh = fakes.B200AdamW(b, lr=1e-3, master_weights=True, state_dtype=torch.float32)
    fresh.load_state_dict(sd)
    for q in b:
        for key in ("exp_avg", "exp_avg_sq", "master"):
            got, want = fresh.state[q][key], ours.state[q][key]
            assert got.dtype == torch.float32 and torch.equal(got, want) and got.data_ptr() != want.data_ptr()
    # a corrupted state (what the un-overridden load_state_dict produced) is refused before any pointer is taken
    fresh.state[b[0]]["master"] = fresh.state[b[0]]["master"].to(torch.bfloat16)
    _set_grads(b, 2)
    with pytest.raises(fakes.B200Error):
        fresh.step()


def test_step_invalidates_the_fused_weight_cache(fakes):
    """The kernels write parameters through raw pointers; modules.fused_weight keys its q|k|v / gate|up concatenations on
    Tensor._version, so step() must bump it or every later forward would run on the pre-training weights."""
    from transformers_b200.modules import fused_weight

    ps = [torch.nn.Parameter(torch.randn(4, 8)) for _ in range(2)]
    holder = torch.nn.Module()
    before = fused_weight(holder, "gate_up", ps).clone()
    assert fused_weight(holder, "gate_up", ps) is fused_weight(holder, "gate_up", ps)  # cached while nothing changes
    opt = fakes.B200AdamW(ps, lr=0.1)
    _set_grads(ps, 1)
    _register(opt)
    v0 = [p._version for p in ps]
    opt.step()
    assert all(p._version > v for p, v in zip(ps, v0))
    after = fused_weight(holder, "gate_up", ps)
    assert not torch.equal(after, before) and torch.equal(after, torch.cat([p.detach() for p in ps]))
    g0 = [p.grad._version for p in ps]
    _fake_ops.register_tensors(*[p.grad for p in ps])
    fakes.clip_grad_norm_(ps, 1e-3)
    assert all(p.grad._version > v for p, v in zip(ps, g0))
