"""GPTConfig — constructor surface and defaults of paddlenlp/transformers/gpt/configuration.py:162-320."""
from ..configuration_utils import PretrainedConfig


class GPTConfig(PretrainedConfig):
    model_type = "gpt"
    attribute_map = {"num_classes": "num_labels", "dropout": "classifier_dropout", "n_positions": "max_position_embeddings",
                     "n_embd": "hidden_size", "n_layer": "num_hidden_layers", "n_head": "num_attention_heads",
                     "n_inner": "intermediate_size", "activation_function": "hidden_activation"}

    def __init__(self, seq_length=1024, vocab_size=50304, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_activation="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=16, initializer_range=0.02, pad_token_id=0, eos_token_id=7,
                 bos_token_id=0, eol_token_id=3, normalize_before=True, scale_qk_coeff=1.0, ignore_index=0, **kwargs):
        self.seq_length = seq_length
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_activation = hidden_activation
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.eol_token_id = eol_token_id
        self.normalize_before = normalize_before
        self.scale_qk_coeff = scale_qk_coeff
        self.ignore_index = ignore_index          # NB: defaults to 0 (gpt/configuration.py:265,300)
        if not normalize_before:
            raise NotImplementedError("post-LN GPT is outside the GPT-2 configuration this build covers")
        super().__init__(pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id,
                         tie_word_embeddings=True, **kwargs)

    @classmethod
    def gpt2_small(cls, **kw):
        """GPT-2-small (124M): BASELINE.json configs[0]."""
        base = dict(vocab_size=50257, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                    max_position_embeddings=1024, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, eos_token_id=50256)
        base.update(kw)
        return cls(**base)
