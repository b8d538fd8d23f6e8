# r02 job 13: two species (test by splitting, bench at cfg-2 size with a second species), CUBEnu order
set -x
python -m pytest tests/test_gpu_two_species.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02n_pytest.log
python bench.py --species 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02n_bench_species2.log 2> gpurun_out/r02n_bench_species2.err; echo "bench rc=$?"
tail -3 gpurun_out/r02n_bench_species2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02n_bench_species2.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], '%.3e'%d['value'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
PY
