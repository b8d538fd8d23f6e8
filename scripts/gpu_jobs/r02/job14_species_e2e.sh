# r02 job 14: two-species deposit with prefetched accumulate; e2e step broken down call by call
set -x
python -m pytest tests/test_gpu_two_species.py -m gpu -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02o_pytest.log
python bench.py --species 2 --steps 5 --warmup 3 --no-cpu --no-late > gpurun_out/r02o_bench_species2.log 2> gpurun_out/r02o_bench_species2.err; echo "bench rc=$?"
tail -3 gpurun_out/r02o_bench_species2.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 > gpurun_out/r02o_bench_e2e.log 2> gpurun_out/r02o_bench_e2e.err; echo "bench rc=$?"
tail -3 gpurun_out/r02o_bench_e2e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02o_bench_species2.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], '%.3e'%d['value'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
d=json.loads(open('gpurun_out/r02o_bench_e2e.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
PY
