# coarse mesh overlapped with the fine mesh on a second stream: full GPU suite + z=49 bench with and without the overlap
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for v in A=1 CUBE_GPU_NO_OVERLAP=1; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
l=json.loads(open("gpurun_out/bench_$v.log").read().strip().splitlines()[-1])
print("$v", round(l["ms_per_step"],2), {k:round(x,2) for k,x in l["phases_ms_per_step"].items()})
PY
done
