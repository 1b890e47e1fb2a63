"""Generate the golden fixtures that pin the oracle (run in the build container; outputs are committed).

    python oracle/make_golden.py            # writes tests/golden/{llama,qwen2}_tiny.pt

PaddlePaddle cannot be imported here, so the fixtures come from the numerical twin the reference itself names:
HuggingFace `transformers` Llama / Qwen2 with transposed Linear weights
(reference test: tests/transformers/llama/test_modeling.py:398-506, LlamaCompatibilityTest).
Each fixture stores the Paddle-layout weights, token ids, HF fp32 logits, the reference criterion's loss on them
(llama/modeling.py:1799-1825; HF's own loss shifts labels and is NOT used) and the HF parameter gradients of that loss.
tests/test_oracle.py checks oracle/llama_ref.py against these on the CPU, with no transformers import needed.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import llama_ref as R  # noqa: E402


def make(model_type: str, path: str):
    import transformers

    cfg = R.RefConfig(vocab_size=160, hidden_size=64, intermediate_size=176, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, rope_theta=500000.0 if model_type == "llama" else 1e6,
                      qkv_bias=(model_type == "qwen2"), model_type=model_type,
                      rms_norm_eps=1e-5 if model_type == "llama" else 1e-6, max_position_embeddings=64)
    w = R.init_weights(cfg, seed=2024, round_bf16=True)
    g = torch.Generator().manual_seed(7)
    for k in w:
        if "norm" in k:
            w[k] = (1.0 + 0.1 * torch.randn(w[k].shape, generator=g)).to(torch.bfloat16).float()
        elif k.endswith("weight"):
            w[k] = (w[k] * 4.0).to(torch.bfloat16).float()        # non-degenerate logits
    cls = transformers.Qwen2ForCausalLM if model_type == "qwen2" else transformers.LlamaForCausalLM
    m = cls(R.hf_config(cfg)).float().eval()
    res = m.load_state_dict(R.to_hf_state_dict(w, cfg), strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    tok = torch.randint(0, cfg.vocab_size, (2, 33), generator=g)
    ids, labels = tok[:, :-1].contiguous(), tok[:, 1:].contiguous()
    labels[1, :5] = -100
    logits = m(input_ids=ids).logits
    loss = R.criterion(logits, labels)
    loss.backward()
    hf_grads = {k: v.grad.clone() for k, v in m.named_parameters()}
    # back to Paddle names / layout
    grads = {}
    for k in w:
        hk = ("model." + k[len(cfg.model_type) + 1:]) if k.startswith(cfg.model_type + ".") else k
        gk = hf_grads[hk]
        grads[k] = gk.t().contiguous() if (k.endswith("_proj.weight") or k == "lm_head.weight") else gk
    torch.save({"config": cfg.__dict__, "weights": {k: v.to(torch.bfloat16) for k, v in w.items()}, "input_ids": ids,
                "labels": labels, "hf_logits_fp32": logits.detach(), "loss_fp32": loss.detach(),
                "grads_fp32": {k: v.to(torch.float32) for k, v in grads.items()},
                "generator": f"oracle/make_golden.py, transformers {transformers.__version__}, torch {torch.__version__}"},
               path)
    print(path, os.path.getsize(path), "bytes; loss", float(loss))


def make_gpt2(path: str):
    """Tiny random HF GPT2LMHeadModel -> reference-named weights, ids, fp32 logits, reference-criterion loss."""
    import transformers

    from oracle import gpt_ref

    torch.manual_seed(11)
    cfg = transformers.GPT2Config(vocab_size=160, n_positions=64, n_embd=64, n_layer=2, n_head=4, resid_pdrop=0.0,
                                  embd_pdrop=0.0, attn_pdrop=0.0, activation_function="gelu_new")
    m = transformers.GPT2LMHeadModel(cfg).float().eval()
    with torch.no_grad():
        for p_ in m.parameters():
            p_.add_(torch.randn_like(p_) * 0.05)
    g = torch.Generator().manual_seed(3)
    tok = torch.randint(1, 160, (2, 33), generator=g)
    ids, labels = tok[:, :-1].contiguous(), tok[:, 1:].contiguous()
    labels[0, :4] = 0                                               # ignore_index = 0
    with torch.no_grad():
        logits = m(input_ids=ids).logits
    w = gpt_ref.from_hf_state_dict(m.state_dict(), 2)
    torch.save({"weights": w, "input_ids": ids, "labels": labels, "hf_logits_fp32": logits,
                "loss_fp32": gpt_ref.criterion(logits, labels), "n_layer": 2, "n_head": 4,
                "generator": f"oracle/make_golden.py, transformers {transformers.__version__}"}, path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    make_gpt2(os.path.join(out, "gpt2_tiny.pt"))
    make("llama", os.path.join(out, "llama_tiny.pt"))
    make("qwen2", os.path.join(out, "qwen2_tiny.pt"))
