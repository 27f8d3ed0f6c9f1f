#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200; every profiled command first runs plainly and must exit 0).
# Outputs go to gpurun_out/; the summaries committed under profiles/ are extracted from them with tools/ncu_extract.py.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-batched --no-prefilter"
$B > $O/p_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_10M.csv $B > $O/p_ncu1.log 2>&1
$B > $O/p_plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemv_topk_kernel -s 6 -c 2 -f -o $O/r02_gemv_10M $B > $O/p_ncu2.log 2>&1
$B > $O/p_plain_bench3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:finalize_published -s 6 -c 2 -f -o $O/r02_exact_cluster_10M $B > $O/p_ncu3.log 2>&1
for MODE in "--int8" ""; do
  TAG=$([ -n "$MODE" ] && echo int8 || echo bf16)
  C="python tools/bench_batch.py --rows 10000000 --steps 1 --warmup 1 $MODE"
  $C > $O/p_plain_batch_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_batch_10M_$TAG.csv $C > $O/p_ncu4_$TAG.log 2>&1
  $C > $O/p_plain_batch2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 1 -f -o $O/r02_gemm2_filter_10M_$TAG $C > $O/p_ncu5_$TAG.log 2>&1
done
C="python tools/bench_batch.py --rows 1000000 --steps 1 --warmup 1 --int8"
$C > $O/p_plain_batch_1M.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_batch_1M_int8.csv $C > $O/p_ncu6.log 2>&1
ls -la $O/*.ncu-rep $O/r02_launches*.csv
