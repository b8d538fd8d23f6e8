# r02 job 22: streamed upload (key pass under the transfer): parity + e2e
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_image.py tests/test_c_driver.py -m gpu -q -x > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02w_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 > gpurun_out/r02w_bench_e2e.log 2> gpurun_out/r02w_bench_e2e.err; echo "bench rc=$?"
tail -3 gpurun_out/r02w_bench_e2e.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 --no-stream-upload > gpurun_out/r02w_bench_e2e_plain_upload.log 2> gpurun_out/r02w_bench_e2e_plain_upload.err; echo "bench rc=$?"
python - <<PY
import json
for f in ('r02w_bench_e2e','r02w_bench_e2e_plain_upload'):
    d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['e2e'])
PY
