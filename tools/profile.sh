#!/bin/bash
# ncu evidence for the round (run under gpurun on ONE GPU).  gpurun_out/ must stay below 64 MiB, so the .ncu-rep files
# are reduced to raw-metric CSVs + a markdown summary (tools/ncu_summary.py) on the box and then deleted.
set -x
mkdir -p gpurun_out
O=gpurun_out
# 1) launch list of the bench command: every kernel with its device time (serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 2600 --csv \
    --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 4 \
    > $O/bench_under_ncu.log 2>&1
# 2) full capture of the tensor-core kernels (GEMM fwd/dX/dW shapes, attention fwd/bwd)
timeout 900 ncu --set full --clock-control none \
    -k regex:'gemm_bf16_kernel|fa_fwd_kernel|fa_bwd_kernel' -s 60 -c 30 -o $O/prof_top \
    python tools/step_probe.py --layers 2 --iters 1 --batch 2 > $O/prof_top.log 2>&1
# 3) full capture of the HBM-bound fusions
timeout 600 ncu --set full --clock-control none \
    -k regex:'rmsnorm|swiglu|rope_kernel|ce_fwd|ce_bwd|embedding|colsum|dq_finish|delta' -s 40 -c 24 -o $O/prof_ew \
    python tools/step_probe.py --layers 2 --iters 1 --batch 2 > $O/prof_ew.log 2>&1
# 4) decode-step kernels
timeout 600 ncu --set full --clock-control none \
    -k regex:'decode_attention|add_rmsnorm|decode_rope|gemm_bf16_kernel|argmax' -s 400 -c 16 -o $O/prof_decode \
    python tools/gen_bench.py --gen 8 --no-graph > $O/prof_decode.log 2>&1
# 5) optimizer kernels on the full 8B flat buffer
timeout 600 ncu --set full --clock-control none -k regex:'adamw|sqnorm' -s 0 -c 3 -o $O/prof_opt \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/prof_opt.log 2>&1
python tools/ncu_summary.py $O/prof_top.ncu-rep $O/prof_ew.ncu-rep $O/prof_decode.ncu-rep $O/prof_opt.ncu-rep > $O/ncu_summary.md
rm -f $O/*.ncu-rep
ls -la $O
