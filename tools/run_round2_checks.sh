# one GPU box: the full GPU suite (request path through the CPython side door), a slice of it through ctypes, the latency breakdown, the bench line
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_gpu_tests9.log; tail -3 gpurun_out/r2_gpu_tests9.log
REBERT_PYCALL=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "vs_oracle or ties or concurrent" 2>&1 | tail -3
python tools/latency_breakdown.py > gpurun_out/r2_latency5.log 2>&1; cat gpurun_out/r2_latency5.log | cut -c1-700
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; tail -c 400 gpurun_out/r2_bench7.err; head -c 300 gpurun_out/r2_bench7.json
