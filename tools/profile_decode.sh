#!/bin/bash
# ncu evidence for the decode-step kernels (run under gpurun on ONE GPU); reports are reduced on the box (64 MiB limit).
set -x
mkdir -p gpurun_out
O=gpurun_out
# 1) launch list of two decode steps at full width (eager launches, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv \
    --log-file $O/launches_decode.csv python tools/gen_bench.py --gen 16 --no-graph > $O/decode_under_ncu.log 2>&1
# 2) full capture of the decode-step kernels (one layer and a half of the chain)
timeout 600 ncu --set full --clock-control none \
    -k regex:'gemm_skinny_kernel|decode_attention_tc|add_rmsnorm|decode_rope|gemm_bf16_kernel|swiglu_fwd' -s 600 -c 14 \
    -o $O/prof_decode2 python tools/gen_bench.py --gen 8 --no-graph > $O/prof_decode2.log 2>&1
python tools/ncu_summary.py $O/prof_decode2.ncu-rep > $O/ncu_decode_summary.md
# SASS evidence: tcgen05 / TMA mnemonics in the new kernels
cuobjdump -sass paddlenlp_b200/build/gemm_skinny.o | grep -oE "UTCHMMA|UTMALDG|UTMAREDG|UTCBAR|UTMAPF|UTMACCTL" | sort | uniq -c > $O/sass_gemm_skinny.txt
cuobjdump -sass paddlenlp_b200/build/decode_attn_tc.o | grep -oE "UTCHMMA|UTMALDG|UTMAREDG|UTCBAR|UTMAPF|LDTM|STTM" | sort | uniq -c > $O/sass_decode_attn_tc.txt
rm -f $O/*.ncu-rep
ls -la $O
