#!/bin/bash
# ncu evidence for the round (run under gpurun on ONE GPU).  Outputs land in gpurun_out/; summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
# 1) launch list of the bench command: every kernel with its device time (serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 2600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 4 \
    > gpurun_out/bench_under_ncu.log 2>&1
# 2) full capture of the tensor-core kernels (GEMM fwd/dX/dW shapes, attention fwd/bwd)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'gemm_bf16_kernel|fa_fwd_kernel|fa_bwd_kernel' -s 60 -c 30 -o gpurun_out/prof_top \
    python tools/step_probe.py --layers 2 --iters 1 --batch 2 > gpurun_out/prof_top.log 2>&1
# 3) full capture of the HBM-bound fusions and the optimizer
timeout 600 ncu --set full --clock-control none \
    -k regex:'rmsnorm|swiglu|rope_kernel|ce_fwd|ce_bwd|embedding|colsum|dq_finish|delta' -s 40 -c 24 -o gpurun_out/prof_ew \
    python tools/step_probe.py --layers 2 --iters 1 --batch 2 > gpurun_out/prof_ew.log 2>&1
# 4) decode-step kernels
timeout 600 ncu --set full --clock-control none \
    -k regex:'decode_attention|add_rmsnorm|decode_rope|gemm_bf16_kernel|argmax' -s 400 -c 16 -o gpurun_out/prof_decode \
    python tools/gen_bench.py --gen 8 --no-graph > gpurun_out/prof_decode.log 2>&1
ls -la gpurun_out
