"""BASELINE.json configs[0]: GPT-2 causal-LM forward + loss on the CPU through the paddlenlp.transformers-style surface
(the reference's own CPU-runnable plumbing case)."""
import math
import os

import torch

from oracle import gpt_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gpt2_tiny.pt")


def test_oracle_and_product_match_hf_golden():
    import paddlenlp_b200.transformers as T

    d = torch.load(GOLD, map_location="cpu", weights_only=False)
    w = d["weights"]
    logits = gpt_ref.forward(d["input_ids"], w, d["n_layer"], d["n_head"])
    assert torch.allclose(logits, d["hf_logits_fp32"], rtol=1e-4, atol=2e-5)            # oracle pinned to HF GPT-2
    assert abs(float(gpt_ref.criterion(logits, d["labels"])) - float(d["loss_fp32"])) < 1e-5
    cfg = T.GPTConfig(vocab_size=160, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                      max_position_embeddings=64, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = T.GPTForCausalLM(cfg)
    missing = model.load_state_dict(w, strict=True)
    loss, out = model(input_ids=d["input_ids"], labels=d["labels"])
    assert torch.allclose(out, d["hf_logits_fp32"], rtol=1e-4, atol=2e-5)
    assert abs(float(loss) - float(d["loss_fp32"])) < 1e-5
    assert model.criterion.ignore_index == 0                                            # gpt/configuration.py:265


def test_config1_gpt2_small_forward_loss_on_cpu():
    """GPT-2-small (124M), batch 2 x seq 128, random init: runs on the CPU in fp32, loss ~ ln(V)."""
    import paddlenlp_b200.transformers as T

    torch.manual_seed(0)
    cfg = T.GPTConfig.gpt2_small()
    model = T.AutoModelForCausalLM.from_config(cfg, dtype="float32")
    n = sum(p.numel() for p in model.parameters())
    assert abs(n / 1e6 - 124.44) < 0.05
    tok = torch.randint(1, cfg.vocab_size, (2, 129))
    with torch.no_grad():
        loss, logits = model(input_ids=tok[:, :-1], labels=tok[:, 1:])
    assert logits.shape == (2, 128, cfg.vocab_size) and logits.dtype == torch.float32
    assert abs(float(loss) - math.log(cfg.vocab_size)) < 0.5
