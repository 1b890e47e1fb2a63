"""AdamW + ClipGradByGlobalNorm on the engine's flat buffers, and the LR schedules the reference's scripts use.

Mirrors what Trainer.create_optimizer builds (paddlenlp/trainer/trainer.py:1717-1750):
    paddle.optimizer.AdamW(learning_rate=lr_scheduler, beta1, beta2, epsilon, parameters, weight_decay,
                           apply_decay_param_fun=<name has no "bias"/"norm">, grad_clip=ClipGradByGlobalNorm(max_grad_norm),
                           multi_precision=True)
One launch pair per step (squared-norm reduction, fused clip+AdamW) over the whole model instead of per-tensor ops.
Schedules: get_scheduler linear/cosine/constant with warm-up (trainer_utils.py) and the pre-training
Linear/CosineAnnealingWithWarmupDecay (llm/run_pretrain.py:520-536).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops


class ClipGradByGlobalNorm:
    def __init__(self, clip_norm: float = 1.0):
        self.clip_norm = float(clip_norm)


class LRScheduler:
    def __init__(self, learning_rate: float):
        self.base_lr = float(learning_rate)
        self.last_epoch = 0

    def get_lr(self) -> float:
        raise NotImplementedError

    def __call__(self) -> float:
        return self.get_lr()

    def step(self):
        self.last_epoch += 1

    def state_dict(self):
        return {"last_epoch": self.last_epoch}

    def set_state_dict(self, sd):
        self.last_epoch = sd["last_epoch"]


class ConstantLR(LRScheduler):
    def get_lr(self):
        return self.base_lr


class ConstantLRWithWarmup(LRScheduler):
    """get_scheduler("constant_with_warmup") (trainer_utils.py get_constant_schedule_with_warmup): linear ramp from 0 over
    `num_warmup_steps`, then the base learning rate."""

    def __init__(self, learning_rate, num_warmup_steps):
        super().__init__(learning_rate)
        self.warmup = int(num_warmup_steps)

    def get_lr(self):
        s = self.last_epoch
        if s < self.warmup:
            return self.base_lr * float(s) / float(max(1, self.warmup))
        return self.base_lr


class LinearDecayWithWarmup(LRScheduler):
    """get_scheduler("linear"): linear warm-up then linear decay to 0 at num_training_steps."""

    def __init__(self, learning_rate, num_training_steps, num_warmup_steps):
        super().__init__(learning_rate)
        self.total, self.warmup = int(num_training_steps), int(num_warmup_steps)

    def get_lr(self):
        s = self.last_epoch
        if s < self.warmup:
            return self.base_lr * s / max(1, self.warmup)
        return self.base_lr * max(0.0, (self.total - s) / max(1, self.total - self.warmup))


class CosineDecayWithWarmup(LRScheduler):
    def __init__(self, learning_rate, num_training_steps, num_warmup_steps, num_cycles: float = 0.5):
        super().__init__(learning_rate)
        self.total, self.warmup, self.cycles = int(num_training_steps), int(num_warmup_steps), num_cycles

    def get_lr(self):
        s = self.last_epoch
        if s < self.warmup:
            return self.base_lr * s / max(1, self.warmup)
        prog = (s - self.warmup) / max(1, self.total - self.warmup)
        return self.base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * self.cycles * 2.0 * prog)))


class LinearAnnealingWithWarmupDecay(LRScheduler):
    """llm/run_pretrain.py:520-536 schedule family: warm-up to max_lr, linear anneal to min_lr over decay_step."""

    def __init__(self, max_lr, min_lr, warmup_step, decay_step, last_epoch=0, verbose=False):
        super().__init__(max_lr)
        self.max_lr, self.min_lr, self.warmup_step, self.decay_step = max_lr, min_lr, warmup_step, decay_step
        self.last_epoch = int(last_epoch)

    def _coeff(self, ratio):
        return 1.0 - ratio

    def get_lr(self):
        s = self.last_epoch
        if self.warmup_step > 0 and s <= self.warmup_step:
            return self.max_lr * s / self.warmup_step
        if s > self.decay_step:
            return self.min_lr
        ratio = (s - self.warmup_step) / max(1, self.decay_step - self.warmup_step)
        return self.min_lr + self._coeff(ratio) * (self.max_lr - self.min_lr)


class CosineAnnealingWithWarmupDecay(LinearAnnealingWithWarmupDecay):
    def _coeff(self, ratio):
        return 0.5 * (math.cos(math.pi * ratio) + 1.0)


def get_scheduler(name, learning_rate, num_warmup_steps=0, num_training_steps=None, num_cycles=0.5, **_):
    name = str(getattr(name, "value", name)).lower()         # SchedulerType enum or plain string
    if name == "linear":
        return LinearDecayWithWarmup(learning_rate, num_training_steps, num_warmup_steps)
    if name == "cosine":
        return CosineDecayWithWarmup(learning_rate, num_training_steps, num_warmup_steps, num_cycles)
    if name == "constant":
        return ConstantLR(learning_rate)
    if name == "constant_with_warmup":
        return ConstantLRWithWarmup(learning_rate, num_warmup_steps)
    raise ValueError(f"unknown lr_scheduler_type {name}")


class AdamW:
    """Flat-buffer AdamW with fp32 master weights (multi_precision) and fused global-norm clipping."""

    def __init__(self, learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, parameters=None, weight_decay=0.01,
                 apply_decay_param_fun=None, grad_clip: Optional[ClipGradByGlobalNorm] = None, multi_precision=True,
                 engine=None):
        if engine is None:
            raise ValueError("AdamW needs the model's DecoderEngine (flat parameter / gradient buffers)")
        if not multi_precision:
            raise NotImplementedError("bf16 parameters are always updated through fp32 master weights (AMP O2)")
        self.engine = engine
        self._lr = learning_rate
        self.beta1, self.beta2, self.eps, self.weight_decay = beta1, beta2, epsilon, weight_decay
        self.grad_clip = grad_clip
        self.step_count = 0
        self.grad_scale = 1.0            # e.g. 1/world_size after a SUM all-reduce
        n = engine.numel
        dev = engine.device
        self.master = torch.empty(n, dtype=torch.float32, device=dev)
        ops.bf16_to_f32(engine.flat_params, self.master)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)

    def get_lr(self) -> float:
        return self._lr() if callable(self._lr) else float(self._lr)

    def set_lr(self, lr):
        self._lr = lr

    def sync_master_from_params(self):
        ops.bf16_to_f32(self.engine.flat_params, self.master)

    def step(self):
        eng = self.engine
        self.step_count += 1
        max_norm = self.grad_clip.clip_norm if self.grad_clip is not None else 0.0
        if max_norm > 0:
            ops.grad_sqnorm(eng.flat_grads, scale=self.grad_scale, out=self.sqnorm)
        ops.adamw_step(eng.flat_params, eng.flat_grads, self.master, self.exp_avg, self.exp_avg_sq,
                       self.sqnorm if max_norm > 0 else None, decay_end=eng.decay_end, lr=self.get_lr(), beta1=self.beta1,
                       beta2=self.beta2, eps=self.eps, weight_decay=self.weight_decay, step=self.step_count,
                       grad_scale=self.grad_scale, max_grad_norm=max_norm)
        eng.params_changed()

    def clear_grad(self, set_to_zero: bool = False):
        self.engine.clear_grad()

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (device scalar; reading it synchronises)."""
        return self.sqnorm.sqrt()

    def state_dict(self):
        return {"step": self.step_count, "master": self.master, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}

    def set_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.master.copy_(sd["master"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])

    # -- per-parameter (reference-named) state: the unified-checkpoint layout -----------------------------------
    # trainer/plugins/unified_checkpoint.py:95-103,437-452: optimizer tensors are keyed "<param name>/moment1_0",
    # "<param name>/moment2_0", "<param name>/beta1_pow_acc_0", "<param name>/beta2_pow_acc_0"; fp32 master weights are
    # keyed by the bare parameter name.
    def named_optimizer_state(self) -> Dict[str, torch.Tensor]:
        eng = self.engine
        m1 = eng.named_views(flat=self.exp_avg)
        m2 = eng.named_views(flat=self.exp_avg_sq)
        # Paddle's accumulators start at beta and are multiplied by beta AFTER each update: after `step` steps they hold
        # beta ** (step + 1) (the value the NEXT step's bias correction uses)
        b1 = torch.tensor([self.beta1 ** (self.step_count + 1)], dtype=torch.float32)
        b2 = torch.tensor([self.beta2 ** (self.step_count + 1)], dtype=torch.float32)
        out: Dict[str, torch.Tensor] = {}
        for k in m1:
            out[k + "/moment1_0"] = m1[k]
            out[k + "/moment2_0"] = m2[k]
            out[k + "/beta1_pow_acc_0"] = b1.clone()      # safetensors refuses aliased tensors
            out[k + "/beta2_pow_acc_0"] = b2.clone()
        return out

    def named_master_weights(self) -> Dict[str, torch.Tensor]:
        return self.engine.named_views(flat=self.master)

    def load_named_state(self, optim_items, master_items, step: Optional[int] = None):
        """`optim_items` / `master_items`: iterables of (key, host tensor) in the layout above.  The step count comes from
        `step` when given, else from beta1_pow_acc (= beta1 ** (step + 1), Paddle's post-update convention)."""
        eng = self.engine
        m1 = eng.named_views(flat=self.exp_avg)
        m2 = eng.named_views(flat=self.exp_avg_sq)
        mw = eng.named_views(flat=self.master)
        seen = set()
        pow_step = None
        with torch.no_grad():
            for key, t in optim_items:
                name, _, kind = key.rpartition("/")
                if kind == "moment1_0":
                    m1[name].copy_(t.to(m1[name].device))
                elif kind == "moment2_0":
                    m2[name].copy_(t.to(m2[name].device))
                elif kind == "beta1_pow_acc_0":
                    v = float(t.reshape(-1)[0])
                    if 0.0 < v < 1.0:
                        pow_step = max(0, round(math.log(v) / math.log(self.beta1)) - 1)
                    elif v == 1.0:
                        pow_step = 0
                    continue
                elif kind == "beta2_pow_acc_0":
                    continue
                else:
                    raise KeyError(f"unknown optimizer state entry {key}")
                seen.add(key)
            for name, t in master_items:
                mw[name].copy_(t.to(mw[name].device))
                seen.add(name)
        missing = [k for k in m1 if (k + "/moment1_0") not in seen or (k + "/moment2_0") not in seen or k not in seen]
        if missing:
            raise KeyError(f"optimizer checkpoint is missing state for {len(missing)} parameters, e.g. {missing[:3]}")
        if step is not None and pow_step is not None and abs(step - pow_step) > 1 and self.beta1 ** step > 1e-30:
            raise ValueError(f"optimizer step {step} disagrees with beta1_pow_acc ({pow_step} steps)")
        self.step_count = int(step if step is not None else (pow_step or 0))
