# r02 job 4: merged brick kick (bulk-copied force bricks) + restructured key/chain pass: parity, A/B at z=49, z=0 state
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_tile.py tests/test_gpu_multi_image.py tests/test_gpu_zz_full_size.py tests/test_gpu_errors.py tests/test_gpu_fft_plans.py -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02d_pytest.log
for v in A=1 CUBE_GPU_OLD_KICK=1 CUBE_GPU_KICK_HOT=16384 CUBE_GPU_KICK_HOT=32768; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02d_bench_$v.log 2> gpurun_out/r02d_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench_$v.log').read().strip().splitlines()[-1])
print('$v', d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
PY
done
python scripts/evolve_bench.py --sweep "CUBE_GPU_OLD_KICK=1" > gpurun_out/r02d_evolve.jsonl 2> gpurun_out/r02d_evolve.err; echo "evolve rc=$?"
tail -3 gpurun_out/r02d_evolve.jsonl | cut -c1-900
