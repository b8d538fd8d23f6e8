# whole half table in shared memory (CUBE_GPU_VT_HOT=32768) A/B at z=49 and z=0; ncu --set full of the particle kernels on the z=0 state of cfg 1
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for v in A=1 CUBE_GPU_VT_HOT=32768; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
l=json.loads(open("gpurun_out/bench_$v.log").read().strip().splitlines()[-1])
print("$v", round(l["ms_per_step"],2), {k:round(x,2) for k,x in l["phases_ms_per_step"].items()})
PY
done
timeout 800 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 560 --sweep "CUBE_GPU_VT_HOT=32768;CUBE_GPU_VT_HOT=28672" > gpurun_out/evolve_cfg2.jsonl 2> gpurun_out/evolve_cfg2.err; echo "evolve2 rc=$?"
grep -v histogram gpurun_out/evolve_cfg2.jsonl | tail -4 | cut -c1-1500; tail -3 gpurun_out/evolve_cfg2.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_fine_deposit|k_drift_count|k_coarse_deposit|k_drift_place_w|k_coarse_kick_w" --launch-skip 385 --launch-count 5 -f -o gpurun_out/prof_z0 \
    python scripts/evolve_bench.py --nc 128 --nnt 2 --max-seconds 300 > gpurun_out/ncu_z0.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_z0.log | cut -c1-300; ls -la gpurun_out/prof_z0.ncu-rep
