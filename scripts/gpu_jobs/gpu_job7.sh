# crowded-cell warp paths: parity tests, uniform-state bench, z=49 -> 0 evolution at cfg 2 (per-phase times of the last window)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_base.log 2> gpurun_out/bench_base.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_base.log | cut -c1-300; grep -o '"phases_ms_per_step".*"cpu' gpurun_out/bench_base.log | cut -c1-900
timeout 700 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 560 > gpurun_out/evolve_cfg2.jsonl 2> gpurun_out/evolve_cfg2.err; echo "evolve2 rc=$?"
tail -2 gpurun_out/evolve_cfg2.jsonl | cut -c1-1200; tail -3 gpurun_out/evolve_cfg2.err
for v in "$@"; do
  env $v timeout 700 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 560 > gpurun_out/evolve_cfg2_$v.jsonl 2> gpurun_out/evolve_cfg2_$v.err; echo "evolve2 $v rc=$?"
  tail -1 gpurun_out/evolve_cfg2_$v.jsonl | cut -c1-1200
done
