#!/bin/bash
# One gpurun session: attention op tests first (decides which forward generation the rest uses), then the full GPU suite,
# the attention micro-benchmark for both generations, and the bench line.   usage: tools/gpu_session.sh <tag> [bench args]
TAG=${1:-r02}; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/${TAG}_smi.log 2>&1
timeout 600 python -m pytest tests/test_ops_gpu.py -q -k "flash_attention" 2>&1 | tail -40 > $OUT/${TAG}_fa_tests.log
if grep -q "failed" $OUT/${TAG}_fa_tests.log; then
  echo "attention tests failed: falling back to forward generation 1 for the rest of this session" >> $OUT/${TAG}_fa_tests.log
  export B200_FA_FWD_IMPL=1
fi
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_ops_gpu.py::test_flash_attention_fwd_bwd \
  --deselect tests/test_ops_gpu.py::test_flash_attention_bench_shapes -rA 2>&1 | grep -v "^PASSED" | tail -120 > $OUT/${TAG}_tests.log
timeout 600 python tools/fa_bench.py > $OUT/${TAG}_fa_bench.log 2>&1
B200_FA_FWD_IMPL=1 timeout 600 python tools/fa_bench.py > $OUT/${TAG}_fa_bench_impl1.log 2>&1
timeout 1500 python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_fa_tests.log; tail -5 $OUT/${TAG}_tests.log; cat $OUT/${TAG}_fa_bench.log | head -3; head -c 1500 $OUT/${TAG}_bench.json
