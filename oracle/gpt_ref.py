"""CPU oracle of the reference's GPT-2 math (TEST INFRASTRUCTURE ONLY — see oracle/llama_ref.py header).

Functional restatement of paddlenlp/transformers/gpt/modeling.py: embeddings :746-759 (word + learned position),
_core_attention :350-385 (q * d^-0.5, triangular mask of finfo.min, softmax), decoder layer :636-700 (pre-LN, eps 1e-5,
tanh-GELU), final norm :455, tied head :1461-1503, criterion :1336-1363 (ignore_index 0, mask = loss > 0).
Pinned against HF GPT2LMHeadModel (same [in,out] Conv1D weight layout) by oracle/make_golden.py -> tests/golden/gpt2_tiny.pt.
"""
import torch
import torch.nn.functional as F


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def forward(input_ids, w, n_layer, n_head):
    """w: dict with the reference parameter names (gpt.embeddings.word_embeddings.weight, gpt.decoder.layers.N.…)."""
    b, s = input_ids.shape
    x = w["gpt.embeddings.word_embeddings.weight"][input_ids] + w["gpt.embeddings.position_embeddings.weight"][torch.arange(s)]
    h = x.shape[-1]
    d = h // n_head
    mask = torch.full((s, s), torch.finfo(torch.float32).min).triu(1)
    for i in range(n_layer):
        p = f"gpt.decoder.layers.{i}."
        n1 = layer_norm(x, w[p + "norm1.weight"], w[p + "norm1.bias"])
        q, k, v = ((n1 @ w[p + f"self_attn.{n}_proj.weight"] + w[p + f"self_attn.{n}_proj.bias"]).view(b, s, n_head, d).transpose(1, 2)
                   for n in ("q", "k", "v"))
        att = torch.softmax((q * d ** -0.5) @ k.transpose(-1, -2) + mask, dim=-1) @ v
        att = att.transpose(1, 2).reshape(b, s, h)
        x = x + att @ w[p + "self_attn.out_proj.weight"] + w[p + "self_attn.out_proj.bias"]
        n2 = layer_norm(x, w[p + "norm2.weight"], w[p + "norm2.bias"])
        m = F.gelu(n2 @ w[p + "linear1.weight"] + w[p + "linear1.bias"], approximate="tanh")
        x = x + m @ w[p + "linear2.weight"] + w[p + "linear2.bias"]
    x = layer_norm(x, w["gpt.decoder.norm.weight"], w["gpt.decoder.norm.bias"])
    return x @ w["gpt.embeddings.word_embeddings.weight"].t()


def criterion(logits, labels, ignore_index=0):
    per = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), labels.reshape(-1), reduction="none", ignore_index=ignore_index)
    mask = (per > 0).float()
    return (per * mask).sum() / mask.sum()


def from_hf_state_dict(sd, n_layer):
    """HF GPT2LMHeadModel -> reference names (c_attn [h,3h] split into q/k/v; Conv1D weights are already [in,out])."""
    w = {"gpt.embeddings.word_embeddings.weight": sd["transformer.wte.weight"],
         "gpt.embeddings.position_embeddings.weight": sd["transformer.wpe.weight"],
         "gpt.decoder.norm.weight": sd["transformer.ln_f.weight"], "gpt.decoder.norm.bias": sd["transformer.ln_f.bias"]}
    for i in range(n_layer):
        hp, p = f"transformer.h.{i}.", f"gpt.decoder.layers.{i}."
        cw, cb = sd[hp + "attn.c_attn.weight"], sd[hp + "attn.c_attn.bias"]
        h = cw.shape[0]
        for j, n in enumerate(("q", "k", "v")):
            w[p + f"self_attn.{n}_proj.weight"] = cw[:, j * h:(j + 1) * h].contiguous()
            w[p + f"self_attn.{n}_proj.bias"] = cb[j * h:(j + 1) * h].contiguous()
        w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"] = sd[hp + "attn.c_proj.weight"], sd[hp + "attn.c_proj.bias"]
        w[p + "linear1.weight"], w[p + "linear1.bias"] = sd[hp + "mlp.c_fc.weight"], sd[hp + "mlp.c_fc.bias"]
        w[p + "linear2.weight"], w[p + "linear2.bias"] = sd[hp + "mlp.c_proj.weight"], sd[hp + "mlp.c_proj.bias"]
        w[p + "norm1.weight"], w[p + "norm1.bias"] = sd[hp + "ln_1.weight"], sd[hp + "ln_1.bias"]
        w[p + "norm2.weight"], w[p + "norm2.bias"] = sd[hp + "ln_2.weight"], sd[hp + "ln_2.bias"]
    return {k: v.clone().float() for k, v in w.items()}
