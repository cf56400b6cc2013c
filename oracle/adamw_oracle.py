"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline may import this; the product never does).

CPU restatement of the optimizer step that follows the hot path.  The reference delegates it to torch,
a third-party dependency that is not vendored under /root/reference (pinned here: torch 2.11.0):
  * Trainer._inner_training_loop -> self.optimizer.step() (trainer.py:1788) with torch.optim.AdamW selected by
    trainer_optimizer.py:201-208;
  * Trainer._clip_grad_norm (trainer.py:2538-2542) -> accelerate -> torch.nn.utils.clip_grad_norm_.
The functions below restate torch's published single-tensor AdamW rule and clip rule; tests/test_optim_cpu.py pins them
against torch.optim.AdamW / torch.nn.utils.clip_grad_norm_ themselves, run in the same process (fp32: to rounding; bf16
parameters: the oracle computes in fp32 on the loaded values and rounds once per stored tensor, like a fused kernel, while
torch's CPU path rounds after every elementary op -- compared within bf16 tolerance)."""
import math


def adamw_step(p, g, m, v, step, lr, beta1, beta2, eps, weight_decay, grad_scale=1.0):
    """One update, returns new (p, m, v) in the dtypes of the inputs; arithmetic in fp32 like the kernel."""
    pf, gf, mf, vf = p.float(), g.float() * grad_scale, m.float(), v.float()
    pf = pf - lr * weight_decay * pf
    mf = mf + (gf - mf) * (1.0 - beta1)
    vf = beta2 * vf + (1.0 - beta2) * gf * gf
    bc1 = 1.0 - beta1 ** step
    bc2_sqrt = math.sqrt(1.0 - beta2 ** step)
    denom = vf.sqrt() / bc2_sqrt + eps
    pf = pf - (lr / bc1) * (mf / denom)
    return pf.to(p.dtype), mf.to(m.dtype), vf.to(v.dtype)


def grad_norm_and_coef(grads, max_norm):
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    coef = min(1.0, max_norm / (total + 1e-6)) if max_norm > 0 else 1.0
    return total, coef
