# r02 job 16: streamed checkpoint with the density deposited once per group and convolved in small batches
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_two_species.py tests/test_gpu_bench_tile.py -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02q_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 > gpurun_out/r02q_bench_e2e.log 2> gpurun_out/r02q_bench_e2e.err; echo "bench rc=$?"
tail -3 gpurun_out/r02q_bench_e2e.err
python bench.py --species 2 --steps 5 --warmup 3 --no-cpu --no-late > gpurun_out/r02q_bench_species2.log 2> gpurun_out/r02q_bench_species2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02q_bench_e2e.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['phases_ms_per_step']['fine_deposit'], d['e2e'])
d=json.loads(open('gpurun_out/r02q_bench_species2.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], '%.3e'%d['value'], d['phases_ms_per_step']['fine_deposit'])
PY
