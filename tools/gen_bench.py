"""Config 5: Llama-3-8B generation decode, 1xB200, batch 64, prompt 128 -> gen 1920 (FusedMultiTransformer KV-cache path).

Reports prefill time, decode tokens/s = B * (gen - 1) / decode time, and the HBM roofline of the decode step
(weights 15.01 GB + KV read 8.39 MB * t per step)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import paddlenlp_b200.transformers as T  # noqa: E402
from paddlenlp_b200.experimental.transformers import LlamaForCausalLMInferenceModel  # noqa: E402


def run(batch=64, prompt=128, gen=1920, layers=0, graph=True, pdl=True, block_attn=False):
    class A:
        pass
    a = A()
    a.batch, a.prompt, a.gen, a.layers, a.no_graph, a.no_pdl, a.block_attn = batch, prompt, gen, layers, not graph, not pdl, block_attn
    return _run(a)


def _run(a):
    cfg = T.LlamaConfig.llama3_8b(num_hidden_layers=a.layers) if a.layers else T.LlamaConfig.llama3_8b()
    m = LlamaForCausalLMInferenceModel(cfg, block_attn=a.block_attn)
    m.init_random(seed=42)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, cfg.vocab_size, (a.batch, a.prompt), generator=g).cuda()
    max_len = a.prompt + a.gen
    caches = m.allocate_caches(a.batch, max_len)
    # warm-up (kernel attributes, allocator)
    m.generate(ids, max_length=min(8, a.gen), eos_token_id=-1, cache_kvs=caches, use_cuda_graph=not a.no_graph, use_pdl=not a.no_pdl)
    torch.cuda.synchronize()
    # prefill alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc = torch.full((a.batch,), a.prompt, dtype=torch.int32, device="cuda")
    e0.record()
    m._prefill(ids, enc, caches)
    e1.record()
    torch.cuda.synchronize()
    prefill_ms = e0.elapsed_time(e1)
    e0.record()
    out, stop, dec = m.generate(ids, max_length=a.gen, eos_token_id=-1, cache_kvs=caches, use_cuda_graph=not a.no_graph,
                                sync_interval=0, use_pdl=not a.no_pdl)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    decode_ms = total_ms - prefill_ms
    steps = a.gen - 1
    L = cfg.num_hidden_layers
    h, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    kvd = cfg.num_key_value_heads * 128
    w_bytes = L * (h * (h + 2 * kvd) + h * h + 3 * h * I) * 2 + V * h * 2
    kv_per_tok = 2 * L * a.batch * kvd * 2
    mean_t = a.prompt + steps / 2.0
    bytes_per_step = w_bytes + kv_per_tok * mean_t
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    ms_step = decode_ms / steps
    achieved = bytes_per_step / (ms_step / 1e3) / 1e9
    rec = dict(workload="Llama-3-8B generation decode, batch 64, prompt 128 -> +1920, FusedMultiTransformer KV-cache path "
                        "(BASELINE.json configs[4])" if (a.batch, a.prompt, a.gen, a.layers) == (64, 128, 1920, 0) else "custom",
               batch=a.batch, prompt=a.prompt, gen=a.gen, layers=L, paged_kv=bool(a.block_attn), prefill_ms=prefill_ms, decode_ms=decode_ms,
               ms_per_step=ms_step,
               decode_tokens_per_s=a.batch * steps / (decode_ms / 1e3), bytes_per_step_gb=bytes_per_step / 1e9,
               achieved_gbs=achieved, hbm_peak_gbs=peaks["hbm_gbs"], roofline_frac=achieved / peaks["hbm_gbs"],
               graph=not a.no_graph, pdl=not a.no_pdl, mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, last_tokens=out[0, -4:].tolist())
    del m, caches
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--prompt", type=int, default=128)
    ap.add_argument("--gen", type=int, default=1920)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--block-attn", action="store_true", help="paged KV cache (FusedBlockMultiTransformer, 64-row blocks)")
    a = ap.parse_args()
    print(json.dumps(_run(a)), flush=True)


if __name__ == "__main__":
    main()
