# r02 job 5: kick staging modes (TMA box / row bulk copies / cp.async), place slimming, timing split of the key+chain pass
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_tile.py tests/test_gpu_multi_image.py -m gpu -q -x > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02e_pytest.log
for v in A=1 CUBE_GPU_KICK_STAGE=1 CUBE_GPU_KICK_STAGE=0 CUBE_GPU_K1_DBG=1 CUBE_GPU_K1_DBG=3 CUBE_GPU_K1_DBG=7; do
  env $v python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02e_bench_$v.log 2> gpurun_out/r02e_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02e_bench_$v.log').read().strip().splitlines()[-1])
    print('$v', d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
except Exception as e:
    print('$v failed', e); print(open('gpurun_out/r02e_bench_$v.err').read()[-600:])
PY
done
