"""Output containers (paddlenlp/transformers/model_outputs.py: CausalLMOutputWithCrossAttentions, BaseModelOutputWithPast)."""
from dataclasses import dataclass
from typing import Any, Optional, Tuple


@dataclass
class BaseModelOutputWithPastAndCrossAttentions:
    last_hidden_state: Any = None
    past_key_values: Optional[Tuple] = None
    hidden_states: Optional[Tuple] = None
    attentions: Optional[Tuple] = None


@dataclass
class CausalLMOutputWithCrossAttentions:
    loss: Any = None
    logits: Any = None
    past_key_values: Optional[Tuple] = None
    hidden_states: Optional[Tuple] = None
    attentions: Optional[Tuple] = None

    def to_tuple(self):
        return tuple(v for v in (self.loss, self.logits, self.past_key_values, self.hidden_states, self.attentions)
                     if v is not None)
