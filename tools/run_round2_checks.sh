# one GPU box: the full GPU suite, the traces DESIGN.md cites, the batched launch list and the bench line
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_gpu_tests7.log; tail -3 gpurun_out/r2_gpu_tests7.log
{
for a in "1250000 10 fused" "1250000 100 fused" "1250000 10 i8 fused" "2264 10 fused" "2269 10 fused dim=32 fp32"; do echo "## trace_gemv.py $a"; python tools/trace_gemv.py $a 2>&1 | cut -c1-1500 | sed -n 2,6p; done
} > gpurun_out/r2_trace7.log 2>&1
grep -A3 "^##" gpurun_out/r2_trace7.log | grep -o '^##.*\|"event_us_per_launch[^}]*\|"cluster_kernel.*' | cut -c1-420
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_batch_1M_int8_c.csv python tools/bench_batch.py --rows 1000000 --int8 --steps 2 --warmup 1 > /dev/null 2>&1
grep "finalize_topk\|select_cand\|gemm2_kernel<1" gpurun_out/r02_launches_batch_1M_int8_c.csv | tail -3 | awk -F'","' '{print $5, $NF}' | cut -c1-120
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; tail -c 400 gpurun_out/r2_bench5.err; head -c 300 gpurun_out/r2_bench5.json
