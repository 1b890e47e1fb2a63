"""Pin the oracle (oracle/llama_ref.py) against the committed golden fixtures (tests/golden/*.pt, produced by
oracle/make_golden.py from HuggingFace transformers — the reference's own declared numerical twin,
tests/transformers/llama/test_modeling.py:398-506) and against properties the reference tests."""
import os

import pytest
import torch

from oracle import llama_ref as R
from oracle import optim_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    d = torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)
    cfg = R.RefConfig(**d["config"])
    w = {k: v.float() for k, v in d["weights"].items()}
    return d, cfg, w


@pytest.mark.parametrize("name", ["llama_tiny.pt", "qwen2_tiny.pt"])
def test_oracle_fp32_matches_hf_golden(name):
    d, cfg, w = load(name)
    logits = R.model_forward(d["input_ids"], w, cfg, mode="fp32")
    # the reference's own compat tolerance is rtol 1e-2 / atol 1e-3; the fp32 restatement is far tighter
    assert torch.allclose(logits, d["hf_logits_fp32"], rtol=1e-4, atol=2e-5)
    loss = R.criterion(logits, d["labels"])
    assert abs(float(loss) - float(d["loss_fp32"])) < 1e-5
    _, _, grads = R.loss_and_grads(d["input_ids"], d["labels"], w, cfg, mode="fp32")
    for k, g in d["grads_fp32"].items():
        assert torch.allclose(grads[k], g, rtol=1e-3, atol=1e-6), k


@pytest.mark.parametrize("name", ["llama_tiny.pt", "qwen2_tiny.pt"])
def test_oracle_bf16_mode_is_a_bf16_perturbation_of_fp32(name):
    d, cfg, w = load(name)
    l32 = R.model_forward(d["input_ids"], w, cfg, mode="fp32")
    l16 = R.model_forward(d["input_ids"], w, cfg, mode="bf16")
    rel = ((l16 - l32).norm() / l32.norm()).item()
    assert 1e-4 < rel < 3e-2          # rounding points are active, and only at bf16 magnitude
    assert torch.equal(l16, l16.to(torch.bfloat16).float())     # outputs are bf16-representable
    # reference tolerance against its twin (LlamaCompatibilityTest): rtol 1e-2, atol 1e-3 on leading logits
    assert torch.allclose(l16[0, 0, :9], d["hf_logits_fp32"][0, 0, :9], rtol=5e-2, atol=5e-2)


def test_criterion_semantics():
    """llama/modeling.py:1799-1825: ignore_index -100, mean over positions whose loss is > 0, 0-count -> sum."""
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(2, 5, 11, generator=g)
    labels = torch.randint(0, 11, (2, 5), generator=g)
    labels[0, :2] = -100
    per = torch.nn.functional.cross_entropy(logits.view(-1, 11), labels.view(-1), reduction="none", ignore_index=-100)
    assert abs(float(R.criterion(logits, labels)) - float(per.sum() / 8)) < 1e-6
    assert float(R.criterion(logits, torch.full((2, 5), -100))) == 0.0


def test_rope_is_rotate_half_convention():
    """llama/modeling.py:557-577 and the naming trap of SURVEY.md §8: training and generation both use rotate-half."""
    d = 8
    cos, sin = R.rope_tables(d, 4, 10000.0)
    x = torch.arange(d, dtype=torch.float32).view(1, 1, 1, d).expand(1, 4, 1, d)
    y = R.apply_rope(x, cos, sin, "fp32")
    p = 3
    inv = 1.0 / (10000.0 ** (torch.arange(0, d, 2).float() / d))
    ang = p * inv
    x1, x2 = x[0, p, 0, : d // 2], x[0, p, 0, d // 2:]
    exp = torch.cat([x1 * ang.cos() - x2 * ang.sin(), x2 * ang.cos() + x1 * ang.sin()])
    assert torch.allclose(y[0, p, 0], exp, atol=1e-5)


def test_kv_cache_consistency_property():
    """tests/transformers/llama/test_modeling.py:171-219: logits of a prefix equal the prefix of the logits (causality),
    the property that makes cached and uncached decoding agree (atol 1e-3 there)."""
    d, cfg, w = load("llama_tiny.pt")
    ids = d["input_ids"][:1]
    full = R.model_forward(ids, w, cfg, mode="fp32")
    pre = R.model_forward(ids[:, :17], w, cfg, mode="fp32")
    assert torch.allclose(full[:, :17], pre, atol=1e-5)


def test_position_ids_default_equals_arange_property():
    d, cfg, w = load("llama_tiny.pt")
    ids = d["input_ids"]
    pos = torch.arange(ids.shape[1]).unsqueeze(0).expand_as(ids)
    a = R.model_forward(ids, w, cfg, mode="fp32")
    b = R.model_forward(ids, w, cfg, mode="fp32", position_ids=pos)
    assert torch.equal(a, b)


def test_init_follows_reference():
    """llama/modeling.py:1386-1436: N(0, 0.02), o_proj / down_proj scaled by 1/sqrt(2L), norm weights 1."""
    cfg = R.RefConfig(vocab_size=64, hidden_size=256, intermediate_size=512, num_hidden_layers=4, num_attention_heads=2,
                      num_key_value_heads=1)
    w = R.init_weights(cfg, seed=1, round_bf16=False)
    assert abs(w["llama.layers.0.self_attn.q_proj.weight"].std().item() - 0.02) < 2e-3
    assert abs(w["llama.layers.0.self_attn.o_proj.weight"].std().item() - 0.02 / (8 ** 0.5)) < 1e-3
    assert abs(w["llama.layers.3.mlp.down_proj.weight"].std().item() - 0.02 / (8 ** 0.5)) < 1e-3
    assert torch.equal(w["llama.norm.weight"], torch.ones(256))


def test_adamw_oracle_against_torch_adamw():
    """Pin the optimizer restatement against torch.optim.AdamW (same decoupled-decay, bias-corrected update rule)."""
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(64, generator=g)
    grads = [torch.randn(64, generator=g) * 0.1 for _ in range(3)]
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    master, m, v = p0.clone(), torch.zeros(64), torch.zeros(64)
    for t, gr in enumerate(grads, 1):
        p.grad = gr.clone()
        opt.step()
        master, m, v, _ = optim_ref.adamw_step(master, m, v, gr, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8,
                                               weight_decay=0.1, step=t, decay_mask=torch.ones(64, dtype=torch.bool),
                                               max_grad_norm=0.0)
    assert torch.allclose(master, p.detach(), rtol=1e-5, atol=1e-6)


def test_flashmask_oracle_packing_invariance():
    """The oracle's FlashMask attention (start rows per key column, fusion_ops.py:218-231): samples packed into one row with
    per-sample position ids give the logits of the samples evaluated alone — the property the CUDA path is then held to."""
    cfg = R.RefConfig(vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=128)
    w = R.init_weights(cfg, seed=3)
    g = torch.Generator().manual_seed(11)
    lens = [23, 1, 40]
    ids = [torch.randint(1, cfg.vocab_size, (n,), generator=g) for n in lens]
    S = 80
    packed = torch.zeros(1, S, dtype=torch.long)
    pos = torch.zeros(1, S, dtype=torch.long)
    ms = torch.zeros(1, S, dtype=torch.int32)
    st = 0
    for t in ids:
        n = len(t)
        packed[0, st:st + n], pos[0, st:st + n], ms[0, st:st + n] = t, torch.arange(n), st + n
        st += n
    ms = torch.maximum(ms, torch.arange(1, S + 1, dtype=torch.int32)[None])      # padding columns: one-token documents
    pre = cfg.model_type
    cos, sin = R.rope_tables(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta)
    x = w[f"{pre}.embed_tokens.weight"][packed]
    for i in range(cfg.num_hidden_layers):
        x = R.decoder_layer(x, w, f"{pre}.layers.{i}.", cfg, cos, sin, "fp32", position_ids=pos, mask_start=ms)
    logits = R.linear(R.rms_norm(x, w[f"{pre}.norm.weight"], cfg.rms_norm_eps, "fp32"), w["lm_head.weight"], None, "fp32")
    st = 0
    for t in ids:
        n = len(t)
        alone = R.model_forward(t[None], w, cfg, mode="fp32")
        assert (logits[0, st:st + n] - alone[0]).abs().max() < 1e-5
        st += n
    assert torch.isfinite(logits).all()                                           # padding rows are finite too


def test_paged_attention_oracle_matches_dense():
    """generation_ref.paged_decode_attention on scattered blocks == plain attention over the gathered cache."""
    import numpy as np
    from oracle import generation_ref as G

    rng = np.random.default_rng(0)
    B, nh, kvh, d, bs, mb = 3, 4, 2, 16, 8, 5
    nb = B * mb + 2
    kc = rng.standard_normal((nb, kvh, bs, d)).astype(np.float32)
    vc = rng.standard_normal((nb, kvh, bs, d)).astype(np.float32)
    tables = rng.permutation(nb)[: B * mb].reshape(B, mb).astype(np.int32)
    lens = np.array([0, 17, mb * bs - 1], np.int32)
    q = rng.standard_normal((B, nh, d)).astype(np.float32)
    out = G.paged_decode_attention(q, kc, vc, tables, lens).reshape(B, nh, d)
    for b in range(B):
        total = min(int(lens[b]) + 1, mb * bs)
        K = np.concatenate([kc[tables[b, j]] for j in range(mb)], axis=1)[:, :total]      # [kvh, total, d]
        V = np.concatenate([vc[tables[b, j]] for j in range(mb)], axis=1)[:, :total]
        for h in range(nh):
            s = K[h // 2] @ q[b, h] / np.sqrt(d)
            p = np.exp(s - s.max())
            assert np.allclose(out[b, h], (p / p.sum()) @ V[h // 2], atol=1e-5)
