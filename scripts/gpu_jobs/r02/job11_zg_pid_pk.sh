# r02 job 11: z-pass arithmetic variants, multi-image particle IDs, device P(k) tests, the z = 0 P(k) gate at 128^3 and 256^3 particles
set -x
python -m pytest tests/test_gpu_multi_image.py tests/test_gpu_power_spectrum.py tests/test_gpu_fft_plans.py tests/test_gpu_bench_tile.py -m gpu -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02l_pytest.log
for v in sc pk mx; do
  CUBE_GPU_ZG=$v python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-late > gpurun_out/r02l_bench_$v.log 2> gpurun_out/r02l_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02l_bench_$v.log').read().strip().splitlines()[-1])
print('$v', d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items() if 'fft' in k})
PY
done
timeout 900 python scripts/pk_gate.py --nc 64 --nnt 2 > gpurun_out/r02l_pk_gate_nc64.json 2> gpurun_out/r02l_pk_gate_nc64.err; echo "pk gate nc64 rc=$?"; cat gpurun_out/r02l_pk_gate_nc64.json | cut -c1-700
timeout 1500 python scripts/pk_gate.py --nc 128 --nnt 2 > gpurun_out/r02l_pk_gate_cfg1.json 2> gpurun_out/r02l_pk_gate_cfg1.err; echo "pk gate cfg1 rc=$?"; cat gpurun_out/r02l_pk_gate_cfg1.json | cut -c1-700
