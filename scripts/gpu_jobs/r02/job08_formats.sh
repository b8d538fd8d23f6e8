# r02 job 8: the four zip formats on the GPU; whole GPU suite; bench
set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02h_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02h_bench.log 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02h_bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
PY
