# one GPU box: the full GPU suite, then the traces and the latency breakdown that DESIGN.md cites
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_gpu_tests6.log; tail -5 gpurun_out/r2_gpu_tests6.log
{
for a in "1250000 10 fused" "1250000 100 fused" "1250000 10 i8 fused" "2264 10 fused" "2269 10 fused dim=32 fp32"; do echo "## trace_gemv.py $a"; python tools/trace_gemv.py $a 2>&1 | cut -c1-1500 | sed -n 2,6p; done
for a in "1250000 10 fused" "2269 10 fused dim=32 fp32"; do echo "## REBERT_FIN_SPLIT=1 trace_gemv.py $a"; REBERT_FIN_SPLIT=1 python tools/trace_gemv.py $a 2>&1 | cut -c1-1500 | sed -n 2,6p; done
} > gpurun_out/r2_trace6.log 2>&1
grep -A3 "^##" gpurun_out/r2_trace6.log | grep -o '^##.*\|"event_us_per_launch[^}]*\|"cluster_kernel.*' | cut -c1-420
python tools/latency_breakdown.py 2269 32 fp32 > gpurun_out/r2_latency3.log 2>&1; cat gpurun_out/r2_latency3.log | cut -c1-600
