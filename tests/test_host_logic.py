"""Host-side logic that needs no GPU: configs, schedules, argument parsing, batch sharding, and the data-parallel
gradient exchange on 2 CPU processes (gloo)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configs_match_reference_defaults_and_presets():
    import paddlenlp_b200.transformers as T

    c = T.LlamaConfig()                      # llama/configuration.py:131-161 defaults (Llama-1-7B)
    assert (c.vocab_size, c.hidden_size, c.intermediate_size, c.num_hidden_layers, c.num_key_value_heads) == (32000, 4096, 11008, 32, 32)
    assert c.rms_norm_eps == 1e-6 and c.rope_theta == 10000.0 and c.use_flash_attention
    assert c.n_embd == 4096                 # attribute_map
    c8 = T.LlamaConfig.llama3_8b()
    assert (c8.vocab_size, c8.intermediate_size, c8.num_key_value_heads, c8.rope_theta) == (128256, 14336, 8, 500000.0)
    q = T.Qwen2Config.qwen2_7b()
    assert (q.hidden_size, q.num_hidden_layers, q.num_attention_heads, q.num_key_value_heads, q.vocab_size) == (3584, 28, 28, 4, 152064)
    rt = T.LlamaConfig.from_dict(json.loads(c8.to_json_string()))
    assert rt.to_dict() == c8.to_dict()
    with pytest.raises(NotImplementedError):
        T.LlamaConfig(tensor_parallel_degree=2)


def test_flop_accounting_matches_survey():
    import paddlenlp_b200.transformers as T
    from paddlenlp_b200.transformers.model_utils import PretrainedModel

    class Dummy(PretrainedModel):
        def __init__(self, cfg):
            torch.nn.Module.__init__(self)
            self.config = cfg

    m = Dummy(T.LlamaConfig.llama3_8b())
    assert abs(m.get_algorithmic_flops_per_token(4096) / 1e9 - 48.249) < 0.01
    assert abs(m.get_model_flops(seq_length=4096) / 4096 / 1e9 - 56.305) < 0.01     # caculate_llm_flops convention
    mq = Dummy(T.Qwen2Config.qwen2_7b())
    assert abs(mq.get_algorithmic_flops_per_token(2048) / 1e9 - 43.655) < 0.01


def test_lr_schedules():
    from paddlenlp_b200.optimizer import CosineAnnealingWithWarmupDecay, LinearAnnealingWithWarmupDecay, get_scheduler

    s = get_scheduler("linear", 1e-3, num_warmup_steps=10, num_training_steps=110)
    vals = []
    for _ in range(111):
        vals.append(s.get_lr()); s.step()
    assert vals[0] == 0.0 and abs(vals[10] - 1e-3) < 1e-12 and abs(vals[60] - 5e-4) < 1e-9 and vals[110] == 0.0
    c = CosineAnnealingWithWarmupDecay(3e-5, 3e-6, warmup_step=30, decay_step=1000)
    c.last_epoch = 30
    assert abs(c.get_lr() - 3e-5) < 1e-12
    c.last_epoch = 1000
    assert abs(c.get_lr() - 3e-6) < 1e-12
    l = LinearAnnealingWithWarmupDecay(1.0, 0.0, warmup_step=0, decay_step=100)
    l.last_epoch = 50
    assert abs(l.get_lr() - 0.5) < 1e-12


def test_argparser_json_and_cmdline(tmp_path, monkeypatch):
    from paddlenlp_b200.trainer import PdArgumentParser, TrainingArguments

    cfg = tmp_path / "a.json"
    cfg.write_text(json.dumps({"per_device_train_batch_size": 1, "gradient_accumulation_steps": 8, "max_steps": 5,
                               "learning_rate": 3e-5, "bf16": True}))
    monkeypatch.setattr(sys, "argv", ["run_pretrain.py", str(cfg), "--max_steps", "7", "--weight_decay", "0.01"])
    (args,) = PdArgumentParser(TrainingArguments).parse_json_file_and_cmd_lines()
    assert args.max_steps == 7 and args.gradient_accumulation_steps == 8 and args.weight_decay == 0.01
    assert args.world_size == 1 and args.data_parallel_degree == 1 and not args.use_hybrid_parallel
    with pytest.raises(NotImplementedError):
        TrainingArguments(sharding="stage2")


def test_shard_rows():
    from paddlenlp_b200.distributed import shard_rows

    assert [shard_rows(64, r, 8) for r in (0, 3, 7)] == [(0, 8), (24, 32), (56, 64)]
    with pytest.raises(ValueError):
        shard_rows(10, 0, 4)


WORKER = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    from paddlenlp_b200 import distributed as D
    D.init_parallel_env("gloo")
    rank, world = D.get_rank(), D.get_world_size()
    assert world == 2

    class FakeEngine:                      # the flat-buffer contract of DecoderEngine, on the CPU
        def __init__(self):
            self.flat_params = torch.full((1024,), float(rank + 1))
            self.flat_grads = torch.arange(1024, dtype=torch.float32) * (rank + 1)
            self.grad_ready_hook = None

    class FakeModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.engine = FakeEngine()

    m = D.DataParallel(FakeModel())
    assert torch.equal(m.engine.flat_params, torch.full((1024,), 1.0)), "rank-0 parameter broadcast at wrap time"
    with m.no_sync():
        assert m._sync is False
    m.sync_gradients()                     # nothing was reduced during backward: ONE all-reduce(SUM) of the whole buffer
    assert torch.equal(m.engine.flat_grads, torch.arange(1024, dtype=torch.float32) * 3)
    mean = m.engine.flat_grads * (1.0 / world)      # the 1/world factor lives in the optimizer's grad_scale
    assert torch.equal(mean, torch.arange(1024, dtype=torch.float32) * 1.5)
    # overlapped exchange: ranges reported final during the last backward are reduced at once, the rest in sync_gradients();
    # every element is summed exactly once
    m.engine.flat_grads = torch.arange(1024, dtype=torch.float32) * (rank + 1)
    with m.no_sync():
        m.prepare_backward()
        assert m.engine.grad_ready_hook is None            # accumulation micro-step: no exchange
    m.prepare_backward()
    hook = m.engine.grad_ready_hook
    hook(900, 1024); hook(300, 640); hook(0, 0)
    assert len(m._pending) == 2
    m.sync_gradients()
    assert m.engine.grad_ready_hook is None and m._pending == []
    assert torch.equal(m.engine.flat_grads, torch.arange(1024, dtype=torch.float32) * 3.0)
    lo, hi = D.shard_rows(16)
    assert (lo, hi) == (rank * 8, rank * 8 + 8)
    torch.distributed.barrier()
    print("OK " + str(rank), flush=True)      # one write per token: the two ranks' lines may interleave only at line ends
""")


def test_data_parallel_exchange_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    import socket

    with socket.socket() as sk:             # a free port (tests/parallel_launch.py in the reference does the same)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK 0" in r.stdout and "OK 1" in r.stdout


def test_zero_padding_dataset_and_collator():
    """Sample packing restated from paddlenlp/datasets/zero_padding_dataset.py and the right-padding collator."""
    import torch

    from paddlenlp_b200.data import DataCollatorForSeq2Seq
    from paddlenlp_b200.datasets import ZeroPaddingMapDataset, generate_greedy_packs

    def rec(n, base):
        return {"input_ids": list(range(base, base + n)), "labels": [-100] * (n // 2) + list(range(n - n // 2)),
                "position_ids": list(range(n)), "attn_mask_startend_row_indices": [n] * n}

    data = [rec(5, 100), rec(4, 200), rec(20, 300), rec(3, 400), rec(6, 500)]       # the 20-token record exceeds max_length
    ds = ZeroPaddingMapDataset(data, tokenizer=None, max_length=10)
    assert len(ds) == 2
    a, b = ds[0], ds[1]
    assert a["input_ids"] == list(range(100, 105)) + list(range(200, 204))          # 5 + 4 fit, + 3 would overflow? no: 5+4+3 > 10
    assert a["attn_mask_startend_row_indices"] == [5] * 5 + [9] * 4                 # each column -> end of its sample
    assert a["position_ids"] == [0, 1, 2, 3, 4, 0, 1, 2, 3]
    assert b["input_ids"] == list(range(400, 403)) + list(range(500, 506))
    assert b["attn_mask_startend_row_indices"] == [3] * 3 + [9] * 6
    packs = generate_greedy_packs([rec(6, 0), rec(5, 10), rec(4, 20), rec(3, 30), rec(2, 40)], 10)
    assert sorted(sorted(len(r["input_ids"]) for r in p) for p in packs) == [[2], [3, 6], [4, 5]]
    g = ZeroPaddingMapDataset(data, max_length=10, greedy_zero_padding=True)
    assert sorted(len(x["input_ids"]) for x in g) == [9, 9]
    batch = DataCollatorForSeq2Seq(max_length=12, pad_token_id=7)([a, b])
    assert batch["input_ids"].shape == (2, 12) and batch["input_ids"][0, -1] == 7 and batch["labels"][0, -1] == -100
    assert batch["attn_mask_startend_row_indices"].dtype == torch.int32
    assert batch["attn_mask_startend_row_indices"][0].tolist() == [5] * 5 + [9] * 4 + [0] * 3


def test_rope_scaling_variants_match_hf_and_oracle():
    """llama/modeling.py:440-554 rotary variants: the product tables (paddlenlp_b200.ops.rope_tables) against HuggingFace's
    ROPE_INIT_FUNCTIONS (independent implementation of the same published formulas) and against the oracle restatement."""
    import transformers
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS

    from oracle import llama_ref as R
    from paddlenlp_b200 import ops
    import paddlenlp_b200.transformers as T

    l3 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}
    hf = transformers.LlamaConfig(hidden_size=4096, num_attention_heads=32, rope_theta=500000.0, max_position_embeddings=131072,
                                  rope_scaling=dict(l3))
    inv_hf, _ = ROPE_INIT_FUNCTIONS["llama3"](hf, "cpu")
    assert torch.equal(ops.rope_inv_freq(128, 500000.0, l3), inv_hf)
    assert torch.equal(R.rope_inv_freq(128, 500000.0, l3), inv_hf)
    # config plumbing: rope_scaling dict (HF Llama-3.1 config.json) and the reference's rope_scaling_type / factor pair
    assert T.LlamaConfig(rope_scaling=dict(l3)).rope_scaling_spec() == l3
    assert T.LlamaConfig(rope_scaling_type="linear", rope_scaling_factor=4.0).rope_scaling_spec() == {"type": "linear", "factor": 4.0}
    assert T.LlamaConfig().rope_scaling_spec() is None
    with pytest.raises(ValueError):
        T.LlamaConfig(rope_scaling_type="yarn")
    # linear: positions / factor (:446-450); HF divides inv_freq instead — same angles up to one fp32 rounding
    cos, sin = ops.rope_tables(128, 64, 10000.0, "cpu", scaling={"type": "linear", "factor": 4.0})
    t = torch.arange(64, dtype=torch.float32)
    want = torch.einsum("i,j->ij", t, ops.rope_inv_freq(128, 10000.0) / 4.0)
    assert (cos - want.cos()).abs().max() < 1e-5 and (sin - want.sin()).abs().max() < 1e-5
    # ntk: base * f^(d/(d-2)) (:467-470)
    assert torch.allclose(ops.rope_inv_freq(128, 10000.0, {"type": "ntk", "factor": 2.0}),
                          ops.rope_inv_freq(128, 10000.0 * 2.0 ** (128 / 126)))
    # dynamic ntk: unchanged up to max_position_embeddings, rescaled base beyond (:482-487) — HF "dynamic" is the same formula
    d = {"type": "dynamic_ntk", "factor": 2.0}
    assert torch.equal(ops.rope_inv_freq(128, 10000.0, d, seq_len=2048, max_position_embeddings=2048), ops.rope_inv_freq(128, 10000.0))
    hfd = transformers.LlamaConfig(hidden_size=4096, num_attention_heads=32, rope_theta=10000.0, max_position_embeddings=2048,
                                   rope_scaling={"rope_type": "dynamic", "factor": 2.0})
    inv_d, _ = ROPE_INIT_FUNCTIONS["dynamic"](hfd, "cpu", seq_len=4096)
    assert torch.allclose(ops.rope_inv_freq(128, 10000.0, d, seq_len=4096, max_position_embeddings=2048), inv_d, rtol=1e-6, atol=0)
    for sc in (l3, {"type": "linear", "factor": 4.0}, {"type": "ntk", "factor": 2.0}):
        a = ops.rope_tables(128, 96, 500000.0, "cpu", scaling=sc)
        b = R.rope_tables(128, 96, 500000.0, scaling=sc)
        assert torch.equal(a[0], b[0][:, :64]) and torch.equal(a[1], b[1][:, :64])


def test_padding_mask_to_flashmask_start_rows():
    """2-D [batch, seq] padding masks (llama/modeling.py:1517-1552) as FlashMask start rows: left padding hides the pad columns
    from every later row; right padding is a no-op under the causal mask; no other pattern is produced."""
    from paddlenlp_b200.transformers.llama.modeling import _mask_rows_from_padding_mask

    m = torch.tensor([[0, 0, 0, 1, 1, 1, 1, 1],      # left padded by 3
                      [1, 1, 1, 1, 1, 0, 0, 0],      # right padded by 3
                      [1, 1, 1, 1, 1, 1, 1, 1]])
    ms = _mask_rows_from_padding_mask(m)
    assert ms.dtype == torch.int32
    assert ms.tolist() == [[1, 2, 3, 8, 8, 8, 8, 8], [8] * 8, [8] * 8]
    # dense check against the reference's expanded mask: row i sees column c iff c <= i and mask[c] (for real rows)
    S = 8
    for b in range(3):
        for i in range(S):
            if not m[b, i]:
                continue
            for c in range(S):
                ref_visible = (c <= i) and bool(m[b, c])
                ours = (c <= i) and (i < int(ms[b, c]))
                assert ref_visible == ours, (b, i, c)


def test_constant_with_warmup_schedule_and_iterable_shard():
    from paddlenlp_b200.optimizer import get_scheduler
    from paddlenlp_b200.trainer.trainer import IterableDatasetShard

    s = get_scheduler("constant_with_warmup", 1e-3, num_warmup_steps=4, num_training_steps=100)
    vals = []
    for _ in range(8):
        vals.append(s.get_lr()); s.step()
    assert vals[:5] == [0.0, 2.5e-4, 5e-4, 7.5e-4, 1e-3] and vals[5:] == [1e-3] * 3      # linear ramp from 0, then constant
    assert get_scheduler("constant", 1e-3, num_warmup_steps=4).get_lr() == 1e-3

    class Stream(torch.utils.data.IterableDataset):
        def __iter__(self):
            return iter(range(22))

    shards = [list(IterableDatasetShard(Stream(), batch_size=2, drop_last=True, num_processes=3, process_index=r)) for r in range(3)]
    assert shards == [[0, 1, 6, 7, 12, 13], [2, 3, 8, 9, 14, 15], [4, 5, 10, 11, 16, 17]]      # disjoint, every world-th batch
    tail = [list(IterableDatasetShard(Stream(), batch_size=2, drop_last=False, num_processes=3, process_index=r)) for r in range(3)]
    assert tail[0][-2:] == [18, 19] and tail[1][-2:] == [20, 21] and tail[2][-2:] == [0, 1]  # partial group completed by wrapping
