# r02 job 10: packed f32x2 complex arithmetic in the line FFTs: microbenchmark of the instructions, parity of every plan, bench; P(k) tests
set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 scripts/microbench/ffma2.cu && /tmp/ffma2 | tee gpurun_out/r02k_ffma2.txt
python -m pytest tests/test_gpu_fft_plans.py tests/test_gpu_bench_tile.py tests/test_gpu_parity.py tests/test_gpu_power_spectrum.py tests/test_gpu_multi_image.py -m gpu -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02k_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-late > gpurun_out/r02k_bench.log 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02k_bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
PY
