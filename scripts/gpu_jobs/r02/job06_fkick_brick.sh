# r02 job 6: non-persistent TMA-staged fine kick (one CTA per brick) against the old L1-gather kernel and the merged persistent one
set -x
CUBE_GPU_OLD_KICK=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_tile.py -m gpu -q -x > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02f_pytest.log
for v in "CUBE_GPU_OLD_KICK=1" "CUBE_GPU_OLD_KICK=1 CUBE_GPU_OLD_FKICK=1"; do
  n=$(echo $v | tr ' =' '__')
  env $v python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02f_bench_$n.log 2> gpurun_out/r02f_bench_$n.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02f_bench_$n.log').read().strip().splitlines()[-1])
    print('$v', d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
except Exception as e:
    print('$v failed', e); print(open('gpurun_out/r02f_bench_$n.err').read()[-600:])
PY
done
