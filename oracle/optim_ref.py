"""CPU oracle of the optimizer step (TEST INFRASTRUCTURE ONLY — see oracle/llama_ref.py header).

Follows the reference's Trainer optimizer set-up (paddlenlp/trainer/trainer.py:1717-1750):
paddle.optimizer.AdamW(beta1, beta2, epsilon, weight_decay, apply_decay_param_fun = "no 'bias'/'norm' in name",
grad_clip = ClipGradByGlobalNorm(max_grad_norm), multi_precision = True) — i.e. fp32 master weights, decoupled
decay applied first, bias-corrected Adam update, bf16 parameter = round(master).  The arithmetic of Paddle's adamw
kernel is not in /root/reference (Paddle core); this restates its documented update rule.
"""
import torch


def clip_coef(grads_fp32: torch.Tensor, max_norm: float) -> torch.Tensor:
    """ClipGradByGlobalNorm: g * max_norm / max(global_norm, max_norm)."""
    norm = grads_fp32.double().pow(2).sum().sqrt().float()
    return max_norm / torch.maximum(norm, torch.tensor(max_norm))


def adamw_step(master, m, v, grad_bf16_as_f32, *, lr, beta1, beta2, eps, weight_decay, step, decay_mask,
               grad_scale=1.0, max_grad_norm=1.0):
    g = grad_bf16_as_f32 * grad_scale
    if max_grad_norm > 0:
        g = g * clip_coef(g, max_grad_norm)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    p = master * torch.where(decay_mask, torch.tensor(1.0 - lr * weight_decay), torch.tensor(1.0))
    denom = v.sqrt() / (1 - beta2 ** step) ** 0.5 + eps
    p = p - (lr / (1 - beta1 ** step)) * (m / denom)
    return p, m, v, p.to(torch.bfloat16)
