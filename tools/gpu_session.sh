#!/bin/bash
# One gpurun session: attention op tests first (they decide which backward generation the rest of the session uses), then the
# full GPU suite, the attention micro-benchmark, and the bench line.   usage: tools/gpu_session.sh <tag> [bench args]
TAG=${1:-r02}; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests/test_ops_gpu.py -q -k "flash_attention" -rf 2>&1 | tail -150 > $OUT/${TAG}_fa_tests.log
if grep -q "failed" $OUT/${TAG}_fa_tests.log; then
  echo "attention tests failed: falling back to generation 1 of the backward for the rest of this session" >> $OUT/${TAG}_fa_tests.log
  export B200_FA_BWD_IMPL=1
fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_ops_gpu.py::test_flash_attention_fwd_bwd \
  --deselect tests/test_ops_gpu.py::test_flash_attention_bench_shapes -rf 2>&1 | grep -v "^loss: \|^PASSED" > $OUT/${TAG}_tests_full.log
tail -60 $OUT/${TAG}_tests_full.log > $OUT/${TAG}_tests.log
if [ "$PARITY_PRINTS" == "1" ]; then
  timeout 900 python -m pytest tests/test_model_gpu.py -q -s -k "two_layer or reorder or bench_shape or full_width or logits_loss" 2>&1 | grep "^\[" > $OUT/${TAG}_parity_prints.log
fi
if [ "$NCU_DECODE" == "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file $OUT/${TAG}_launches_decode.csv \
    python tools/gen_bench.py --gen 16 > $OUT/${TAG}_decode_ncu.log 2>&1
  timeout 600 python tools/gen_bench.py > $OUT/${TAG}_gen_bench.log 2>&1
fi
timeout 600 python tools/fa_bench.py > $OUT/${TAG}_fa_bench.log 2>&1
if [ "$NCU_FA" == "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fa_fwd2_kernel|fa_bwd2_kernel|fa_bwd_kernel" -s 2 -c 2 \
    -o $OUT/${TAG}_fa_prof -f python tools/fa_probe.py 2 > $OUT/${TAG}_fa_prof.log 2>&1
fi
if [ "$NCU_GEMM" == "1" ]; then
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 \
    --csv --log-file $OUT/${TAG}_gemm_traffic.csv python tools/gemm_shapes.py > $OUT/${TAG}_gemm_traffic.log 2>&1
fi
if [ "$SKIP_BENCH" != "1" ]; then
  timeout 1500 python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
fi
tail -c 1500 $OUT/${TAG}_fa_tests.log; tail -12 $OUT/${TAG}_tests.log; head -3 $OUT/${TAG}_fa_bench.log; head -c 600 $OUT/${TAG}_bench.json
