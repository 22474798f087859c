"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the optimizer tail that follows the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

Follows, line by line:
  binary_seg/utils/utils.py:7-17     clip_gradient: param.grad.data.clamp_(-grad_clip, grad_clip)
  binary_seg/MyTrain_med.py:85-86    clip_gradient(optimizer, opt.clip); optimizer.step()
  binary_seg/MyTrain_med.py:148-149  torch.optim.Adam(params, opt.lr)       (betas (0.9, 0.999), eps 1e-8, weight_decay 0)
  EMCAD/trainer.py:86                optim.AdamW(model.parameters(), lr=base_lr, weight_decay=0.0001)
The arithmetic itself lives in a third-party dependency absent from /root/reference: torch.optim (reference pin
torch==2.0.1, pranet2.yaml:138; installed 2.11).  Its published single-tensor algorithm (torch/optim/adam.py
_single_tensor_adam, adamw.py) is restated here in float32 numpy; tests/test_oracle_golden.py pins this restatement
against the installed torch.optim.Adam / AdamW run on CPU.
"""
import numpy as np

F32 = np.float32


def clamp_adam_step(p, g, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, clip=0.5,
                    grad_scale=1.0):
    """One optimizer step on float32 arrays (any shape); returns (p, m, v) after step number `step` (1-based)."""
    p, g, m, v = (np.asarray(a, dtype=F32) for a in (p, g, m, v))
    b1, b2 = F32(betas[0]), F32(betas[1])
    g = np.clip(g * F32(grad_scale), F32(-clip), F32(clip)) if clip else g * F32(grad_scale)   # utils.py:7-17 (after the DP mean)
    if decoupled:
        p = p * F32(1.0 - lr * weight_decay)              # adamw.py: param.mul_(1 - lr * weight_decay)
    elif weight_decay != 0.0:
        g = g + F32(weight_decay) * p                     # adam.py: grad = grad.add(param, alpha=weight_decay)
    m = m + (g - m) * (F32(1.0) - b1)                     # exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + (F32(1.0) - b2) * g * g                  # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1.0 - float(betas[0]) ** step
    bc2 = 1.0 - float(betas[1]) ** step
    step_size = F32(lr / bc1)
    denom = np.sqrt(v) / F32(np.sqrt(bc2)) + F32(eps)     # (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - step_size * (m / denom)                       # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p.astype(F32), m.astype(F32), v.astype(F32)
