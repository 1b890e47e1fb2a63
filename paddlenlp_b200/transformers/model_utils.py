"""PretrainedModel base: the slice of paddlenlp/transformers/model_utils.py (:921 class, :1101 from_config-style
construction, :1140 recompute_enable) that the decoder hot path exercises.  Parameters are torch Parameters that
alias the engine's flat bf16 buffer; `.grad` aliases the flat gradient buffer."""
from __future__ import annotations

import json
import os
from typing import Dict

import torch
from torch import nn

from .decoder_engine import DecoderEngine


class _CausalLMLossFn(torch.autograd.Function):
    """Bridges `loss.backward()` (trainer.py:2243) to the engine's explicit backward."""

    @staticmethod
    def forward(ctx, anchor, engine, input_ids, labels, position_ids, ignore_index):
        loss_out, logits = engine.forward_loss(input_ids, labels, position_ids, ignore_index)
        ctx.engine = engine
        ctx.mark_non_differentiable(logits)
        return loss_out[0].clone(), logits

    @staticmethod
    def backward(ctx, gloss, _glogits):
        g = gloss.detach().to(torch.float32).reshape(1).contiguous()
        ctx.engine.backward(1.0, g)          # upstream scale stays on the device: no host sync
        return None, None, None, None, None, None


class PretrainedModel(nn.Module):
    config_class = None
    base_model_prefix = ""

    def __init__(self, config):
        super().__init__()
        self.config = config

    # -- construction -------------------------------------------------------------------------------
    def _build_engine(self, config, device=None):
        self.engine = DecoderEngine(config, device=device, prefix=self.base_model_prefix)
        self._anchor = nn.Parameter(torch.zeros(1, device=self.engine.device), requires_grad=True)
        self._named = {}
        grads = self.engine.named_views(grads=True)
        for name, view in self.engine.named_views().items():
            prm = nn.Parameter(view, requires_grad=True)
            prm.grad = grads[name]
            self._named[name] = prm
        seed = getattr(config, "seed", 42)
        self.engine.init_weights(seed)

    @classmethod
    def from_config(cls, config, dtype: str = "bfloat16", **kwargs):
        if dtype not in ("bfloat16", torch.bfloat16):
            raise NotImplementedError("the hot path computes in bf16 (AMP O2: parameters are created in bf16)")
        return cls(config, **kwargs)

    _from_config = from_config

    @classmethod
    def from_pretrained(cls, path, config=None, dtype="bfloat16", **kwargs):
        """Local directory with config.json (+ optional model_state.pt).  Hub download is out of scope."""
        if config is None:
            config = cls.config_class.from_pretrained(path)
        model = cls(config, **kwargs)
        wfile = os.path.join(path, "model_state.pt") if os.path.isdir(path) else None
        if wfile and os.path.exists(wfile):
            model.set_state_dict(torch.load(wfile, map_location="cpu"))
        return model

    def save_pretrained(self, save_directory: str):
        os.makedirs(save_directory, exist_ok=True)
        self.config.save_pretrained(save_directory)
        torch.save({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()},
                   os.path.join(save_directory, "model_state.pt"))

    # -- parameters ---------------------------------------------------------------------------------
    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        for k, v in self._named.items():
            yield (prefix + k, v)

    def parameters(self, recurse: bool = True):
        for _, v in self.named_parameters():
            yield v

    def state_dict(self, *args, **kwargs) -> Dict[str, torch.Tensor]:
        return {k: v.detach() for k, v in self._named.items()}

    def set_state_dict(self, sd: Dict[str, torch.Tensor]):
        self.engine.load_state_dict(sd)

    load_state_dict = set_state_dict

    def num_parameters(self) -> int:
        return self.engine.num_parameters()

    def recompute_enable(self):
        # model_utils.py:1140.  Activation memory is bounded by micro-batching here (gradient accumulation into the
        # flat buffer) rather than by recomputation; accepted for API compatibility.
        self.config.recompute = True

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # -- FLOP accounting ----------------------------------------------------------------------------
    def get_model_flops(self, batch_size=1, seq_length=None, **kwargs):
        """caculate_llm_flops (paddlenlp/transformers/utils.py:963-1003): MHA-sized, non-causal convention,
        x3 for fwd+bwd.  Kept for the reference's `interval_hardware_tflops_per_device` log key."""
        c = self.config
        s = seq_length or getattr(c, "seq_length", 2048)
        h, L, V, I = c.hidden_size, c.num_hidden_layers, c.vocab_size, c.intermediate_size
        flops_per_layer = 2 * s * h * h * 4 + 2 * s * s * h * 2 + 2 * s * h * I * 3
        return 3 * batch_size * (L * flops_per_layer + 2 * s * h * V)

    def get_algorithmic_flops_per_token(self, seq_length=None) -> float:
        """Honest count: GQA-sized projections, causal attention, x3 for fwd+bwd."""
        c = self.config
        s = seq_length or getattr(c, "seq_length", 2048)
        h, L, V, I = c.hidden_size, c.num_hidden_layers, c.vocab_size, c.intermediate_size
        kvd = c.num_key_value_heads * (h // c.num_attention_heads)
        per_layer = 2 * (h * h + 2 * h * kvd + h * h + 3 * h * I) + 2 * s * h
        return 3.0 * (L * per_layer + 2 * h * V)
