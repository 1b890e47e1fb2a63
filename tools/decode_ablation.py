"""Where does the decode step go?  Replays the 32-layer decode kernel chain (CUDA graph + PDL, as generate() does) with
kernel families removed one at a time; the difference to the full chain is that family's marginal cost in the pipeline."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paddlenlp_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
B, h, I, nh, kvh, d, L, max_len = 64, 4096, 14336, 32, 8, 128, 32, 2048
BF = torch.bfloat16


def main():
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.02).to(BF)
    qkv_w = [rnd((nh + 2 * kvh) * d, h) for _ in range(L)]
    o_w = [rnd(nh * d, h) for _ in range(L)]
    f1_w = [rnd(h, 2 * I) for _ in range(L)]
    f2_w = [rnd(I, h) for _ in range(L)]
    ln = torch.ones(h, dtype=BF, device=dev)
    caches = [torch.zeros(2, B, kvh, max_len, d, dtype=BF, device=dev) for _ in range(L)]
    cos, sin = ops.rope_tables(d, max_len, 500000.0, dev)
    x0 = rnd(B, h)
    fix_qkv = rnd(B, (nh + 2 * kvh) * d)
    fix_attn = rnd(B, nh * d)
    fix_act = rnd(B, I)
    fix_ln = rnd(B, h)

    def chain(skip, lens):
        residual = x0
        ln_out, _ = ops.add_rmsnorm(x0, None, ln, 1e-5, want_residual=False)
        for i in range(L):
            if "gemm_qkv" in skip:
                qkv = fix_qkv
            else:
                acc = ops.gemm_skinny_f32(ln_out, qkv_w[i], trans_b=True, tag="splitk_qkv")
                qkv = fix_qkv if "rope" in skip else ops.decode_rope_append_f32(acc, None, caches[i], cos, sin, lens, nh, kvh, d)
            attn = fix_attn if "attn" in skip else ops.decode_attention(qkv, caches[i], lens, nh, kvh, d)
            if "gemm_o" in skip:
                pass
            else:
                acc = ops.gemm_skinny_f32(attn, o_w[i], tag="splitk_h")
                if "norm" not in skip:
                    ln_out, residual = ops.add_rmsnorm_f32(acc, residual, ln, 1e-5)
            if "gemm_f1" in skip:
                act = fix_act
            else:
                if os.environ.get("B200_FFN1", "plain") == "fused":
                    act = ops.gemm_swiglu_skinny(ln_out if "norm" not in skip else fix_ln, f1_w[i])
                elif os.environ.get("B200_FFN1", "plain") == "epi":
                    _, act = ops.gemm_swiglu(ln_out if "norm" not in skip else fix_ln, f1_w[i], cta_group=1, store_gate_up=False)
                elif os.environ.get("B200_FFN1", "plain") == "skinny":
                    acc1 = ops.gemm_skinny_f32(ln_out if "norm" not in skip else fix_ln, f1_w[i], tag="splitk_ffn1")
                    act = fix_act if "swiglu" in skip else ops.swiglu_fwd_f32(acc1)
                else:
                    ffn1 = ops.gemm(ln_out if "norm" not in skip else fix_ln, f1_w[i], cta_group=1)
                    act = fix_act if "swiglu" in skip else ops.swiglu_fwd(ffn1)
            if "gemm_f2" not in skip:
                acc = ops.gemm_skinny_f32(act, f2_w[i], tag="splitk_h")
                if "norm" not in skip:
                    ln_out, residual = ops.add_rmsnorm_f32(acc, residual, ln, 1e-5)
        return residual

    lib = _lib.load()
    lib.b200_set_skinny_gemm(int(os.environ.get("B200_SKINNY", "1")))
    results = {}
    for t in (128, 1024, 2040):
        lens = torch.full((B,), t, dtype=torch.int32, device=dev)
        for name, skip in (("full", ()), ("no_attn", ("attn",)), ("no_norm", ("norm",)), ("no_rope", ("rope",)),
                           ("no_swiglu", ("swiglu",)), ("no_small", ("attn", "norm", "rope", "swiglu")),
                           ("only_qkv", ("attn", "norm", "rope", "swiglu", "gemm_o", "gemm_f1", "gemm_f2")),
                           ("only_o", ("attn", "norm", "rope", "swiglu", "gemm_qkv", "gemm_f1", "gemm_f2")),
                           ("only_f1", ("attn", "norm", "rope", "swiglu", "gemm_qkv", "gemm_o", "gemm_f2")),
                           ("only_f2", ("attn", "norm", "rope", "swiglu", "gemm_qkv", "gemm_o", "gemm_f1")),
                           ("only_attn", ("norm", "rope", "swiglu", "gemm_qkv", "gemm_o", "gemm_f1", "gemm_f2"))):
            if t != 128 and name not in ("full", "no_attn", "only_attn"):
                continue
            for pdl in (1, 0):
                lib.b200_set_pdl(pdl)
                chain(skip, lens)                       # warm-up (workspaces, attributes)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    with torch.cuda.graph(gr, stream=s):
                        chain(skip, lens)
                for _ in range(3):
                    gr.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
                results[f"t{t}.{name}.pdl{pdl}"] = round(e0.elapsed_time(e1) / 10, 4)
                print(json.dumps({"t": t, "variant": name, "pdl": pdl, "ms": results[f"t{t}.{name}.pdl{pdl}"]}), flush=True)
    lib.b200_set_pdl(0)


if __name__ == "__main__":
    main()
