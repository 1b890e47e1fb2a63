"""A/B of the fused GEMM epilogues at the bench shapes (one 8192-token micro-batch of Llama-3-8B):
   gate|up + SwiGLU (mode 4) vs GEMM -> swiglu_fwd;  down-proj dX + SwiGLU backward (mode 5) vs GEMM -> swiglu_bwd;
   rmsnorm fwd / bwd.  Each variant is timed as a back-to-back chain over rotating buffers larger than L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paddlenlp_b200 import ops  # noqa: E402

dev = "cuda:0"
BF = torch.bfloat16
T, h, I = 8192, 4096, 14336
NBUF = 3


def timed(fn, iters=12, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s, sc=0.05: (torch.randn(*s, device=dev, generator=g) * sc).to(BF)
    x = [rnd(T, h, sc=1.0) for _ in range(NBUF)]
    w_gu = rnd(h, 2 * I)
    w_dn = rnd(I, h)
    gu = [rnd(T, 2 * I, sc=1.0) for _ in range(NBUF)]
    dy = [rnd(T, h, sc=1.0) for _ in range(NBUF)]
    out_gu = torch.empty(T, 2 * I, dtype=BF, device=dev)
    out_m = torch.empty(T, I, dtype=BF, device=dev)
    dm = torch.empty(T, I, dtype=BF, device=dev)
    dgu = torch.empty(T, 2 * I, dtype=BF, device=dev)
    res = {}
    res["fwd_unfused_ms"] = timed(lambda i: ops.swiglu_fwd(ops.gemm(x[i % NBUF], w_gu, out=out_gu), out=out_m))
    res["fwd_fused_ms"] = timed(lambda i: ops.gemm_swiglu(x[i % NBUF], w_gu, gate_up=out_gu, out=out_m))
    res["bwd_gemm_only_ms"] = timed(lambda i: ops.gemm(dy[i % NBUF], w_dn, out=dm, trans_b=True))
    res["bwd_unfused_ms"] = timed(lambda i: ops.swiglu_bwd(gu[i % NBUF], ops.gemm(dy[i % NBUF], w_dn, out=dm, trans_b=True), dgate_up=dgu))
    res["bwd_fused_ms"] = timed(lambda i: ops.gemm_swiglu_bwd(dy[i % NBUF], w_dn, gu[i % NBUF], dgate_up=dgu))
    wn = torch.ones(h, dtype=BF, device=dev)
    y, rstd = ops.rmsnorm_fwd(x[0], wn, 1e-5)
    res["rmsnorm_fwd_us"] = 1e3 * timed(lambda i: ops.rmsnorm_fwd(x[i % NBUF], wn, 1e-5), iters=50)
    dwn = torch.zeros(h, dtype=BF, device=dev)
    res["rmsnorm_bwd_us"] = 1e3 * timed(lambda i: ops.rmsnorm_bwd(dy[i % NBUF], x[i % NBUF], wn, rstd, dwn, dres=dy[(i + 1) % NBUF]), iters=50)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
