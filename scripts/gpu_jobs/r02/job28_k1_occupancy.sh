# r02 job 28: key+chain kernel compiled for five CTAs per SM (64 -> 48 registers)
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_zip_formats.py -m gpu -q -x > gpurun_out/r02ze_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02ze_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-cfg1 --no-e2e > gpurun_out/r02ze_bench.log 2> gpurun_out/r02ze_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ze_bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['phases_ms_per_step']['drift_key'], 'late', d['late_time']['ms_per_step'], d['late_time']['phases_ms_per_step']['drift_key'])
PY
