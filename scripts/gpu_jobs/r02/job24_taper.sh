# r02 job 24: tapering batches in the streamed step (16,16,16,8,4,4 tiles), two species on two images
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_image.py tests/test_gpu_two_species.py -m gpu -q -x > gpurun_out/r02za_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02za_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 > gpurun_out/r02za_bench_e2e.log 2> gpurun_out/r02za_bench_e2e.err; echo "bench rc=$?"
tail -3 gpurun_out/r02za_bench_e2e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02za_bench_e2e.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'])
PY
