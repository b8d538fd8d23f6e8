# r02 job 32: fine kick compiled for eight CTAs per SM (40 -> 32 registers, 112 bytes of spills)
set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drift_then_kicks or full_steps" > gpurun_out/r02zi_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02zi_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 --no-e2e > gpurun_out/r02zi_bench.log 2> gpurun_out/r02zi_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zi_bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['phases_ms_per_step']['fine_kick'])
PY
