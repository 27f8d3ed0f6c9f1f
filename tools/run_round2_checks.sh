# one GPU box: the full GPU suite, then the exact-pass kernel's selection variants side by side (traces + end-to-end latency)
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_gpu_tests10.log; tail -3 gpurun_out/r2_gpu_tests10.log
run() { if [ "$1" = default ]; then shift; env -u REBERT_FIN_SPLIT "$@"; else shift; env REBERT_FIN_SPLIT=1 "$@"; fi; }
{
for v in default split; do for a in "1250000 10 fused" "2264 10 fused" "2269 10 fused dim=32 fp32"; do echo "## $v: trace_gemv.py $a"; run $v python tools/trace_gemv.py $a 2>&1 | cut -c1-1500 | sed -n 2,6p; done; done
} > gpurun_out/r2_trace8.log 2>&1
grep -A3 "^##" gpurun_out/r2_trace8.log | grep -o '^##.*\|"event_us_per_launch[^}]*\|"cluster_kernel.*' | cut -c1-330
for v in default split; do echo "== $v"; run $v python tools/latency_breakdown.py 2269 32 fp32 | cut -c1-420; run $v python tools/latency_breakdown.py 1250000 1536 bf16 | cut -c1-420; done 2>&1 | tee gpurun_out/r2_latency6.log
