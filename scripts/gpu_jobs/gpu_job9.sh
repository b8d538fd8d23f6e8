# crowded-cell paths (dense-brick fixed-point fine deposit, REDUX coarse deposit, streamed drift count): parity tests,
# z=49 bench under a few settings, z=0 state of cfg 2 under several thresholds
set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for v in A=1 CUBE_GPU_DENSE_DEPOSIT=0 CUBE_GPU_COUNT_MINB=8 CUBE_GPU_HEAVY_COUNT=24; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
l=json.loads(open("gpurun_out/bench_$v.log").read().strip().splitlines()[-1])
print("$v", round(l["ms_per_step"],2), {k:round(x,2) for k,x in l["phases_ms_per_step"].items() if k in ("drift_count","fine_deposit","coarse_deposit")})
PY
done
timeout 800 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 560 --sweep "$1" > gpurun_out/evolve_cfg2.jsonl 2> gpurun_out/evolve_cfg2.err; echo "evolve2 rc=$?"
tail -14 gpurun_out/evolve_cfg2.jsonl | grep -v histogram | cut -c1-1500; tail -3 gpurun_out/evolve_cfg2.err
