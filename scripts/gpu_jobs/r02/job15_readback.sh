# r02 job 15: small read-backs through mapped memory (no copy-engine queueing behind the streamed checkpoint): tests + e2e
set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02p_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 > gpurun_out/r02p_bench_e2e.log 2> gpurun_out/r02p_bench_e2e.err; echo "bench rc=$?"
tail -3 gpurun_out/r02p_bench_e2e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02p_bench_e2e.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
PY
