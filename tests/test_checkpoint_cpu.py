"""CPU tests of the weight-file plumbing: sharded safetensors layout, HF <-> Paddle conversion pinned to
the oracle's independent name map, TrainerState JSON, checkpoint discovery.  No kernels run: the model is only a container
for its flat parameter buffer here."""
import json
import os

import pytest
import torch

from oracle import llama_ref as R
from paddlenlp_b200.transformers import conversion_utils as cu


def test_plan_shards_and_sizes():
    assert cu.parse_size("5GB") == 5 * 10 ** 9 and cu.parse_size("1MiB") == 1 << 20 and cu.parse_size(123) == 123
    with pytest.raises(ValueError):
        cu.parse_size("lots")
    plan = cu.plan_shards([("a", 40), ("b", 40), ("c", 100), ("d", 10), ("e", 10)], 64)
    assert plan == [["a"], ["b"], ["c"], ["d", "e"]]              # oversize tensor gets its own shard, order kept


def test_sharded_roundtrip_and_index(tmp_path):
    g = torch.Generator().manual_seed(0)
    sd = {f"t{i}": torch.randn(8, 16, generator=g).to(torch.bfloat16) for i in range(5)}
    sd["noncontig"] = torch.randn(16, 8, generator=g).t()
    files = cu.save_sharded(sd, str(tmp_path), max_shard_size=600)
    idx = json.load(open(tmp_path / cu.SAFE_WEIGHTS_INDEX_NAME))
    assert idx["metadata"]["total_size"] == sum(v.numel() * v.element_size() for v in sd.values())
    assert set(idx["weight_map"]) == set(sd) and all(f in files for f in idx["weight_map"].values())
    assert all(f.startswith("model-0000") and "-of-0000" in f for f in idx["weight_map"].values())
    back = cu.load_sharded(str(tmp_path))
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    # re-saving with a different shard count removes the stale files
    cu.save_sharded(sd, str(tmp_path), max_shard_size="1GB")
    assert sorted(os.listdir(tmp_path)) == [cu.SAFE_WEIGHTS_NAME]
    assert all(torch.equal(v, sd[k]) for k, v in cu.iter_sharded(str(tmp_path)))
    os.remove(tmp_path / cu.SAFE_WEIGHTS_NAME)
    with pytest.raises(FileNotFoundError):
        cu.load_sharded(str(tmp_path))


@pytest.mark.parametrize("model_type", ["llama", "qwen2"])
def test_hf_conversion_matches_oracle_name_map(model_type):
    cfg = R.RefConfig(vocab_size=64, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, qkv_bias=(model_type == "qwen2"), model_type=model_type)
    w = R.init_weights(cfg, seed=3)
    hf = R.to_hf_state_dict(w, cfg)                               # oracle's independent statement of modeling.py:1243-1274
    mine = cu.paddle_to_hf_state_dict(w, model_type)
    assert set(mine) == set(hf) and all(torch.equal(mine[k], hf[k]) for k in hf)
    back = cu.hf_to_paddle_state_dict({**hf, "model.layers.0.self_attn.rotary_emb.inv_freq": torch.ones(4)}, model_type)
    assert set(back) == set(w) and all(torch.equal(back[k], w[k]) for k in w)
    assert cu.looks_like_hf(hf) and not cu.looks_like_hf(w)


@pytest.mark.parametrize("hf_format", [False, True])
def test_model_save_and_from_pretrained(tmp_path, hf_format):
    import paddlenlp_b200.transformers as T

    cfg = T.Qwen2Config(vocab_size=64, hidden_size=256, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                        num_key_value_heads=1, max_position_embeddings=64)
    m = T.Qwen2ForCausalLM(cfg, device="cpu")
    with torch.no_grad():
        for i in range(2):
            m.engine.p[f"l{i}.qkv_b"].normal_()                    # biases are zero-initialised: make them matter
    m.save_pretrained(str(tmp_path), max_shard_size="200KB", hf_format=hf_format)
    assert os.path.isfile(tmp_path / "config.json") and os.path.isfile(tmp_path / cu.SAFE_WEIGHTS_INDEX_NAME)
    keys = set(json.load(open(tmp_path / cu.SAFE_WEIGHTS_INDEX_NAME))["weight_map"])
    assert ("model.layers.0.self_attn.q_proj.bias" in keys) == hf_format
    assert ("qwen2.layers.0.self_attn.q_proj.bias" in keys) != hf_format
    m2 = T.Qwen2ForCausalLM.from_pretrained(str(tmp_path), device="cpu")
    assert m2.config.hidden_size == 256 and torch.equal(m2.engine.flat_params, m.engine.flat_params)
    # a checkpoint with a tensor missing is refused
    victim = sorted(set(json.load(open(tmp_path / cu.SAFE_WEIGHTS_INDEX_NAME))["weight_map"].values()))[0]
    from safetensors.torch import load_file, save_file
    t = load_file(str(tmp_path / victim))
    t.pop(sorted(t)[0])
    save_file(t, str(tmp_path / victim))
    with pytest.raises(KeyError):
        T.Qwen2ForCausalLM.from_pretrained(str(tmp_path), device="cpu")


def test_trainer_state_json_and_last_checkpoint(tmp_path):
    from paddlenlp_b200.trainer.trainer import TrainerState, get_last_checkpoint

    st = TrainerState(global_step=7, epoch=1.5, max_steps=10, log_history=[{"loss": 1.0, "global_step": 7}])
    st.save_to_json(str(tmp_path / "trainer_state.json"))
    raw = json.load(open(tmp_path / "trainer_state.json"))
    assert raw["global_step"] == 7 and raw["log_history"][0]["loss"] == 1.0 and "best_model_checkpoint" in raw
    assert TrainerState.load_from_json(str(tmp_path / "trainer_state.json")) == st
    assert get_last_checkpoint(str(tmp_path)) is None
    for n in (2, 10, 9):
        os.makedirs(tmp_path / f"checkpoint-{n}")
    os.makedirs(tmp_path / "checkpoint-11.tmp")
    assert get_last_checkpoint(str(tmp_path)).endswith("checkpoint-10")


def test_gradient_ready_ranges_tile_the_flat_buffer():
    """The ranges the engine reports to the data-parallel exchange during backward (lm_head, each layer's matrices, embedding,
    vector tail) cover the flat gradient buffer exactly once."""
    import paddlenlp_b200.transformers as T

    cfg = T.Qwen2Config(vocab_size=64, hidden_size=256, intermediate_size=72, num_hidden_layers=3, num_attention_heads=2,
                        num_key_value_heads=1, max_position_embeddings=64)
    eng = T.Qwen2ForCausalLM(cfg, device="cpu").engine
    rs = [eng._range("head", "head")] + [eng._range(f"l{i}.qkv_w", f"l{i}.down_w") for i in range(3)]
    rs += [eng._range("embed", "embed"), (eng.decay_end, eng.numel)]
    pos = 0
    for lo, hi in sorted(rs):
        assert lo == pos and hi > lo
        pos = hi
    assert pos == eng.numel
