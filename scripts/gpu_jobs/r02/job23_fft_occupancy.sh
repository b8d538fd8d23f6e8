# r02 job 23: occupancy of the two issue/barrier-bound FFT passes: x inverse at 3 CTAs per SM (40 registers), z pass single-buffered at 3 CTAs per SM
set -x
for zg in sc sb; do
  CUBE_GPU_ZG=$zg python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 --no-e2e > gpurun_out/r02y_bench_$zg.log 2> gpurun_out/r02y_bench_$zg.err; echo "bench $zg rc=$?"
done
CUBE_GPU_ZG=sb python -m pytest tests/test_gpu_fft_plans.py tests/test_gpu_bench_tile.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02y_pytest_sb.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02y_pytest_sb.log
python - <<PY
import json
for f in ('sc','sb'):
    d=json.loads(open('gpurun_out/r02y_bench_%s.log'%f).read().strip().splitlines()[-1])
    p=d['phases_ms_per_step']
    print(f, d['ms_per_step'], {k: round(p[k],2) for k in ('fine_fft_x','fine_fft_y','fine_fft_z_green','fine_ifft_y','fine_ifft_x')})
PY
