# r02 job 2: region fine deposit (fixed-point shared-memory atomics): parity + brick-shape A/B
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_tile.py tests/test_gpu_multi_image.py tests/test_gpu_zz_full_size.py -m gpu -q -x > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02b_pytest.log
for b in 884 888 844; do
  CUBE_GPU_FD_BRICK=$b python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02b_bench_$b.log 2> gpurun_out/r02b_bench_$b.err; echo "bench $b rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02b_bench_$b.log').read().strip().splitlines()[-1])
print($b, d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
PY
done
