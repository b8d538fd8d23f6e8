# error-path tests + GPU suite + z=49 bench + z=0 evolve (micro-optimisations of the key pass and the deposit walk)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_base.log 2> gpurun_out/bench_base.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_base.log").read().strip().splitlines()[-1])
print(round(l["ms_per_step"],2), {k:round(x,2) for k,x in l["phases_ms_per_step"].items()})
PY
timeout 800 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 560 > gpurun_out/evolve_cfg2.jsonl 2> gpurun_out/evolve_cfg2.err; echo "evolve2 rc=$?"
grep -v histogram gpurun_out/evolve_cfg2.jsonl | tail -2 | cut -c1-1500; tail -3 gpurun_out/evolve_cfg2.err
