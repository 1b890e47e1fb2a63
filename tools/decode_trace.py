"""Kernel timeline of the decode step INSIDE the CUDA graph (torch.profiler / CUPTI activity records, so kernels keep their
programmatic-dependent-launch overlap — ncu would serialise them).  Generates prompt -> +gen tokens and aggregates the kernel
records of the last `--steps` decode steps: per kernel name the launches per step, mean duration, and the step's wall span."""
import argparse
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import paddlenlp_b200.transformers as T  # noqa: E402
from paddlenlp_b200.experimental.transformers import LlamaForCausalLMInferenceModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--prompt", type=int, default=1024)
    ap.add_argument("--gen", type=int, default=24)
    ap.add_argument("--timeline", action="store_true", help="also print start / end (us) of every kernel of two middle layers of the last step")
    a = ap.parse_args()
    cfg = T.LlamaConfig.llama3_8b()
    m = LlamaForCausalLMInferenceModel(cfg)
    m.init_random(seed=42)
    ids = torch.randint(0, cfg.vocab_size, (a.batch, a.prompt), generator=torch.Generator().manual_seed(1)).cuda()
    caches = m.allocate_caches(a.batch, a.prompt + a.gen + 8)
    m.generate(ids, max_length=8, eos_token_id=-1, cache_kvs=caches, use_cuda_graph=True, use_pdl=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.generate(ids, max_length=a.gen, eos_token_id=-1, cache_kvs=caches, use_cuda_graph=True, sync_interval=0, use_pdl=True)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    evs = [e for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    evs.sort(key=lambda e: e.time_range.start)
    # decode steps: split at the embedding kernel of each step
    starts = [i for i, e in enumerate(evs) if "embedding_fwd" in e.name]
    steps = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
    steps = steps[-8:]                                   # the last full steps (context ~ prompt + gen)
    agg = collections.OrderedDict()
    spans, busy = [], []
    for lo, hi in steps:
        seg = evs[lo:hi]
        spans.append(seg[-1].time_range.end - seg[0].time_range.start)
        # union of kernel intervals = time at least one kernel is running
        t_end, b = seg[0].time_range.start, 0.0
        for e in seg:
            s0, s1 = max(e.time_range.start, t_end), e.time_range.end
            if s1 > s0:
                b += s1 - s0
                t_end = s1
        busy.append(b)
        prev_end = seg[0].time_range.start
        names = [e.name.split("(")[0][-60:] for e in seg]
        mult = {k: max(1, round(names.count(k) / cfg.num_hidden_layers)) for k in set(names)}
        seen = collections.Counter()
        for i, e in enumerate(seg):
            k = names[i]
            if mult[k] > 1:                              # same kernel at several places of a layer (o-proj / ffn2, the two norms)
                idx = seen[k] % mult[k]
                seen[k] += 1
                k = f"{k} #{idx}"
            # the GEMMs of a layer differ only by shape: tag them with their position in the layer's kernel sequence
            v = agg.setdefault(k, [0, 0.0, 0.0])
            v[0] += 1
            v[1] += e.time_range.end - e.time_range.start
            # exclusive share: with programmatic dependent launch a kernel is resident (parked in griddepcontrol.wait) long
            # before its predecessor ends, so its own duration overlaps; the distance between consecutive END times is what it
            # adds to the step
            v[2] += max(0.0, e.time_range.end - prev_end)
            prev_end = max(prev_end, e.time_range.end)
    timeline = None
    if a.timeline:
        lo, hi = steps[-1]
        seg = evs[lo:hi]
        per_layer = max(1, (len(seg) - 4) // cfg.num_hidden_layers)
        first = 1 + 15 * per_layer                      # kernel 0 is the embedding, then per_layer kernels per layer
        t0 = seg[first].time_range.start
        timeline = [{"kernel": e.name.split("(")[0][-48:], "start_us": round(e.time_range.start - t0, 2),
                     "end_us": round(e.time_range.end - t0, 2)} for e in seg[first:first + 2 * per_layer]]
    n = len(steps)
    out = {"context": a.prompt + a.gen, "steps": n, "step_span_us": sum(spans) / n, "gpu_busy_us": sum(busy) / n,
           "kernels": {k: {"per_step": v[0] / n, "mean_us": v[1] / v[0], "sum_us_per_step": v[1] / n,
                           "exclusive_us_per_step": v[2] / n, "exclusive_us_per_launch": v[2] / v[0]} for k, v in agg.items()}}
    out["sum_of_kernel_durations_us"] = sum(v["sum_us_per_step"] for v in out["kernels"].values())
    if timeline is not None:
        out["timeline_two_layers"] = timeline
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
