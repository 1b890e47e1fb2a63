"""Stand-alone streaming rate of the decode-step kernels (run under gpurun).

Every probe replays ONE kernel family over ROT different weight / cache buffers (more bytes than the 126 MB L2, so each launch streams from
HBM) inside a CUDA graph, and reports microseconds per launch and the HBM rate of the bytes the launch must read.  Unlike
tools/decode_ablation.py (marginal cost inside the 32-layer chain) this isolates a kernel from its neighbours: it answers "is the
kernel bound by HBM, by the number of SMs that own a tile, or by the bytes in flight per SM?" """
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paddlenlp_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
BF = torch.bfloat16
ROT = 8


def graph_time(fns, reps=6):
    """fns: list of callables (one launch each).  Returns us per launch of the whole list replayed as one CUDA graph."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fns:
                f()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))


def rnd(*shape):
    return (torch.randn(*shape, device=dev) * 0.02).to(BF)


def main():
    which = set(sys.argv[1:]) or {"gemm", "attn"}
    M = 64
    x4096, x14336 = rnd(M, 4096), rnd(M, 14336)
    if "gemm" in which:
        for inter in (14336, 18944):
            ws = [rnd(4096, 2 * inter) for _ in range(ROT)]
            act = torch.empty(M, inter, dtype=BF, device=dev)
            out = torch.empty(M, 2 * inter, dtype=BF, device=dev)
            for name, fn in (("persistent gemm [64,4096]x[4096,2I]", lambda w: ops.gemm(x4096, w, out=out, cta_group=1)),
                             ("persistent gemm + SwiGLU epilogue (ffn1)",
                              lambda w: ops.gemm_swiglu(x4096, w, out=act, cta_group=1, store_gate_up=False)),
                             ("skinny gemm + SwiGLU epilogue (ffn1)", lambda w: ops.gemm_swiglu_skinny(x4096, w, out=act))):
                us = graph_time([(lambda w=w: fn(w)) for w in ws])
                print(json.dumps(dict(probe=name, inter=inter, tiles=inter // 128, us=us, tb_s=4096 * 2 * inter * 2 / us / 1e6)),
                      flush=True)
            del ws
        # swapped-operand split-K kernel (two CTAs per SM)
        for name, K, N, trans_b in (("qkv", 4096, 6144, True), ("o_proj", 4096, 4096, False), ("ffn2", 14336, 4096, False),
                                    ("ffn1 shape", 4096, 28672, False)):
            ws = [rnd(N, K) if trans_b else rnd(K, N) for _ in range(ROT)]
            x = x14336 if K == 14336 else x4096
            for split_k in ((0,) if name != "ffn2" else (0, 7, 9)):
                fns = []
                for i, w in enumerate(ws):
                    fns.append(lambda w=w, i=i: ops.gemm_skinny_f32(x, w, trans_b=trans_b, split_k=split_k, tag=f"probe{i}"))
                us = graph_time(fns)
                print(json.dumps(dict(probe=f"skinny gemm {name}", K=K, N=N, split_k=split_k, us=us, tb_s=K * N * 2 / us / 1e6)),
                      flush=True)
            del ws
    if "chain" in which:
        # the six kernels between two attention calls of a layer (out-linear .. next qkv) against ops.decode_layer_chain
        h, inter, qn = 4096, 14336, 6144
        Ws = [dict(o=rnd(h, h), f1=rnd(h, 2 * inter), f2=rnd(inter, h), q=rnd(qn, h)) for _ in range(ROT)]
        ln_w = torch.ones(h, dtype=BF, device=dev)
        attn, res = rnd(M, h), rnd(M, h)

        def unfused(w):
            acc = ops.gemm_skinny_f32(attn, w["o"], tag="p_h")
            ln, r = ops.add_rmsnorm_f32(acc, res, ln_w, 1e-5)
            act = ops.gemm_swiglu_skinny(ln, w["f1"])
            acc = ops.gemm_skinny_f32(act, w["f2"], tag="p_h")
            ln, r = ops.add_rmsnorm_f32(acc, r, ln_w, 1e-5)
            ops.gemm_skinny_f32(ln, w["q"], trans_b=True, tag="p_q")

        def chained(w):
            ops.decode_layer_chain(attn, w["o"], ln_w, w["f1"], w["f2"], ln_w, w["q"], res, 1e-5, qkv_tag="p_q2")

        lib = _lib.load()
        for pdl in (0, 1):
            lib.b200_set_pdl(pdl)
            us_u = graph_time([(lambda w=w: unfused(w)) for w in Ws])
            us_c = graph_time([(lambda w=w: chained(w)) for w in Ws])
            print(json.dumps(dict(probe="layer chain: 6 kernels vs one persistent kernel", pdl=pdl, unfused_us=us_u, chained_us=us_c,
                                  weights_mb=(h * h + 3 * h * inter + qn * h) * 2 / 1e6)), flush=True)
        lib.b200_set_pdl(0)
        # per-phase stamps of one launch
        n_cta = 2 * torch.cuda.get_device_properties(0).multi_processor_count
        stamps = torch.zeros(n_cta, 6, 2, dtype=torch.int64, device=dev)
        lib.b200_decode_layer_chain_debug(stamps.data_ptr())
        for w in Ws[:3]:
            chained(w)
        torch.cuda.synchronize()
        lib.b200_decode_layer_chain_debug(None)
        st = stamps.cpu().double()
        t0 = st[st > 0].min()
        rows = []
        for ph in range(6):
            rdy, pub = st[:, ph, 0], st[:, ph, 1]
            rows.append(dict(phase=ph, ready_first_us=float((rdy[rdy > 0].min() - t0) / 1e3) if (rdy > 0).any() else None,
                             ready_last_us=float((rdy.max() - t0) / 1e3) if (rdy > 0).any() else None,
                             pub_first_us=float((pub[pub > 0].min() - t0) / 1e3), pub_last_us=float((pub.max() - t0) / 1e3)))
        print(json.dumps(dict(probe="layer chain phase stamps (us from the first stamp; ready = inputs seen by a producer, pub = CTA "
                                    "published the phase)", phases=rows)), flush=True)
    if "attn" in which:
        B, nh, kvh, d, max_len = 64, 32, 8, 128, 2048
        caches = [torch.randn(2, B, kvh, max_len, d, device=dev).to(BF) for _ in range(ROT)]
        qkv = rnd(B, (nh + 2 * kvh) * d)
        out = torch.empty(B, nh * d, dtype=BF, device=dev)
        for ctx in (1048, 1900):
            lens = torch.full((B,), ctx - 1, dtype=torch.int32, device=dev)
            for ns in (1, 2, 3, 5, 9, 0):
                try:
                    us = graph_time([(lambda c=c: ops.decode_attention(qkv, c, lens, nh, kvh, d, out=out, num_splits=ns)) for c in caches])
                except Exception as e:  # a split count the library rejects
                    print(json.dumps(dict(probe="decode attention", ctx=ctx, num_splits=ns, error=str(e)[:120])), flush=True)
                    continue
                byts = 2 * B * kvh * ctx * d * 2
                print(json.dumps(dict(probe="decode attention", ctx=ctx, num_splits=ns, us=us, tb_s=byts / us / 1e6)), flush=True)


if __name__ == "__main__":
    main()
