"""Qwen2 modeling classes — paddlenlp/transformers/qwen2/modeling.py surface (same decoder block as Llama plus
q/k/v bias :478-480, eps 1e-6, criterion :1149-1181) on the native engine."""
from ..llama.modeling import LlamaForCausalLM, LlamaModel, LlamaPretrainingCriterion
from ..model_utils import PretrainedModel
from .configuration import Qwen2Config

__all__ = ["Qwen2Model", "Qwen2PretrainedModel", "Qwen2ForCausalLM", "Qwen2PretrainingCriterion"]


class Qwen2PretrainedModel(PretrainedModel):
    config_class = Qwen2Config
    base_model_prefix = "qwen2"


class Qwen2PretrainingCriterion(LlamaPretrainingCriterion):
    """loss[loss > 0].mean() (qwen2/modeling.py:1177-1179) — the same masked mean as the Llama criterion."""


class Qwen2Model(LlamaModel):
    config_class = Qwen2Config
    base_model_prefix = "qwen2"


class Qwen2ForCausalLM(LlamaForCausalLM):
    config_class = Qwen2Config
    base_model_prefix = "qwen2"
