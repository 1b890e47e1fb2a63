"""paddlenlp.transformers surface kept by this build (SURVEY.md §1 "public interface we keep")."""
from ..optimizer import CosineAnnealingWithWarmupDecay, LinearAnnealingWithWarmupDecay
from .configuration_utils import LlmMetaConfig, PretrainedConfig, llmmetaclass
from .gpt.configuration import GPTConfig
from .gpt.modeling import GPTForCausalLM, GPTLMHeadModel, GPTModel, GPTPretrainingCriterion
from .llama.configuration import LlamaConfig
from .llama.modeling import LlamaForCausalLM, LlamaModel, LlamaPretrainedModel, LlamaPretrainingCriterion
from .model_outputs import CausalLMOutputWithCrossAttentions
from .qwen2.configuration import Qwen2Config
from .qwen2.modeling import Qwen2ForCausalLM, Qwen2Model, Qwen2PretrainedModel, Qwen2PretrainingCriterion

_CONFIGS = {"llama": LlamaConfig, "qwen2": Qwen2Config, "gpt": GPTConfig}
_CAUSAL_LM = {"llama": LlamaForCausalLM, "qwen2": Qwen2ForCausalLM, "gpt": GPTForCausalLM}


class AutoConfig:
    """paddlenlp/transformers/auto/configuration.py — local config.json only."""

    @staticmethod
    def from_pretrained(path, **kwargs):
        import json
        import os

        f = os.path.join(path, "config.json") if os.path.isdir(path) else path
        with open(f) as fh:
            d = json.load(fh)
        return _CONFIGS[d.get("model_type", "llama")].from_dict(d, **kwargs)


class AutoModelForCausalLM:
    """paddlenlp/transformers/auto/modeling.py:71 — dispatch on config.model_type."""

    @staticmethod
    def from_config(config, dtype="bfloat16", **kwargs):
        return _CAUSAL_LM[config.model_type].from_config(config, dtype=dtype, **kwargs)

    @staticmethod
    def from_pretrained(path, **kwargs):
        cfg = AutoConfig.from_pretrained(path)
        return _CAUSAL_LM[cfg.model_type].from_pretrained(path, config=cfg, **kwargs)


class _OutOfScope:
    """Names the reference's training scripts import but that lie outside the data-parallel decoder hot path: importing them
    works (so the scripts' import block is unchanged), using them raises with the reason."""
    _why = ""

    def __init__(self, *a, **kw):
        raise NotImplementedError(self._why)

    @classmethod
    def from_pretrained(cls, *a, **kw):
        raise NotImplementedError(cls._why)

    from_config = from_pretrained


class AutoTokenizer(_OutOfScope):
    _why = ("tokenizers are outside this build's scope (SURVEY.md §2.1): feed token ids; a tokenizer object is only carried "
            "through Trainer(tokenizer=...) untouched")


class AutoModelForCausalLMPipe(_OutOfScope):
    _why = "pipeline parallelism: pure data-parallel replication only (pipeline_parallel_degree must be 1)"


def register_sequence_parallel_allreduce_hooks(model, accumulation_steps, fuse_sequence_parallel_allreduce):
    raise NotImplementedError("sequence parallelism needs tensor_parallel_degree > 1: pure data parallelism only")
