"""Run every GEMM shape of the Llama-3-8B training step once (for an ncu capture of DRAM traffic per shape).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 \
        --csv --log-file gpurun_out/gemm_traffic.csv python tools/gemm_shapes.py
    python tools/gemm_shapes.py --summarise gpurun_out/gemm_traffic.csv > profiles/r02_gemm_traffic.json   # here, no GPU needed

M = 8192 tokens (micro-batch 2 x 4096).  Order = the order printed by --list; each shape runs ONCE after the L2 has been flushed
by a 192 MB write, so the DRAM bytes are those of a cold L2 (as in the step, where 1.5 GB of other traffic separates two uses)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

T, h, I, V, QKV = 8192, 4096, 14336, 128256, 6144
# (name, M, N, K, trans_a, trans_b, epilogue)   — shapes as the engine issues them (decoder_engine.py)
SHAPES = [
    ("qkv fwd", T, QKV, h, 0, 0, "none"), ("o fwd (+residual)", T, h, h, 0, 0, "residual"),
    ("gate|up fwd (+swiglu)", T, 2 * I, h, 0, 0, "swiglu"), ("down fwd (+residual)", T, h, I, 0, 0, "residual"),
    ("lm_head fwd", T, V, h, 0, 0, "none"),
    ("lm_head dX", T, h, V, 0, 1, "none"), ("lm_head dW", h, V, T, 1, 0, "none"),
    ("down dX (+swiglu bwd)", T, I, h, 0, 1, "swiglu_bwd"), ("down dW", I, h, T, 1, 0, "none"),
    ("gate|up dX", T, h, 2 * I, 0, 1, "none"), ("gate|up dW", h, 2 * I, T, 1, 0, "none"),
    ("o dX", T, h, h, 0, 1, "none"), ("o dW", h, h, T, 1, 0, "none"),
    ("qkv dX", T, h, QKV, 0, 1, "none"), ("qkv dW", h, QKV, T, 1, 0, "none"),
]


def algorithmic_bytes(M, N, K, epi):
    b = 2 * (M * K + K * N + M * N)
    if epi == "residual":
        b += 2 * M * N
    if epi == "swiglu":
        b += 2 * M * (N // 2)
    if epi == "swiglu_bwd":                 # N = I: reads gate|up (2 M I), writes d(gate|up) (2 M I) instead of d(m) (M I)
        b += 2 * M * N * 2 + 2 * M * N
    return b


def run():
    import torch

    from paddlenlp_b200 import ops

    dev = "cuda"
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
    for name, M, N, K, ta, tb, epi in SHAPES:
        a = torch.randn((K, M) if ta else (M, K), device=dev).to(torch.bfloat16)
        b = torch.randn((N, K) if tb else (K, N), device=dev).to(torch.bfloat16) * 0.02
        flush.zero_()
        if epi == "swiglu":
            ops.gemm_swiglu(a, b)
        elif epi == "swiglu_bwd":
            gu = torch.randn(M, 2 * N, device=dev).to(torch.bfloat16)
            flush.zero_()
            ops.gemm_swiglu_bwd(a, b, gu)
        elif epi == "residual":
            r = torch.randn(M, N, device=dev).to(torch.bfloat16)
            flush.zero_()
            ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), residual=r)
        else:
            ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb))
        torch.cuda.synchronize()
        del a, b
    print("ok")


def summarise(csv_path):
    import csv

    rows = []
    with open(csv_path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    per = {}
    for r in rd:
        if "gemm_bf16_kernel" not in r.get("Kernel Name", ""):
            continue
        key = r["ID"]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6,
                "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1)
        per.setdefault(key, {})[r["Metric Name"]] = val * mult
    ids = sorted(per, key=lambda k: int(k))
    out = []
    for (name, M, N, K, ta, tb, epi), k in zip(SHAPES, ids):
        m = per[k]
        dram = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        alg = algorithmic_bytes(M, N, K, epi)
        out.append({"gemm": name, "M": M, "N": N, "K": K, "epilogue": epi, "dram_bytes": dram, "algorithmic_bytes": alg,
                    "ratio": round(dram / alg, 3), "ms": round(m.get("gpu__time_duration.sum", 0), 4),
                    "tflops": round(2.0 * M * N * K / (m.get("gpu__time_duration.sum", 1) / 1e3) / 1e12, 1)})
    print(json.dumps({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum (cold L2, one launch per "
                                "shape, tools/gemm_shapes.py)", "shapes": out}, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        run()
