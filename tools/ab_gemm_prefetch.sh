set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -3
for pf in 1 0; do
  for rows in 1000000 10000000; do
    REBERT_GEMM_L2_PREFETCH=$pf python tools/bench_batch.py --rows $rows --int8 --steps 3 --warmup 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PF=$pf rows=$rows int8 ms', d['ms_per_batch'])"
  done
  REBERT_GEMM_L2_PREFETCH=$pf python tools/bench_batch.py --rows 1000000 --steps 3 --warmup 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PF=$pf rows=1M bf16 ms', d['ms_per_batch'])"
  REBERT_GEMM_L2_PREFETCH=$pf ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm2 -c 6 --csv --log-file gpurun_out/r02_ab_pf${pf}_10M_int8.csv python tools/bench_batch.py --rows 10000000 --int8 --steps 2 --warmup 1 > /dev/null 2>&1
  grep gemm2 gpurun_out/r02_ab_pf${pf}_10M_int8.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
done
