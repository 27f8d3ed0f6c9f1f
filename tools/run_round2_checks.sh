# one GPU box: the full GPU suite, the latency breakdown and the bench line
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_gpu_tests8.log; tail -3 gpurun_out/r2_gpu_tests8.log
python tools/latency_breakdown.py > gpurun_out/r2_latency4.log 2>&1; cat gpurun_out/r2_latency4.log | cut -c1-700
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; tail -c 400 gpurun_out/r2_bench6.err; head -c 300 gpurun_out/r2_bench6.json
