# A/B of opt-in kernel variants: tests once, then bench.py per environment setting given as arguments ("VAR=1" ...)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_base.log 2> gpurun_out/bench_base.err; echo "bench rc=$?"
for v in "$@"; do
  env $v python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_$v.log 2>&1; echo "pytest $v rc=$?"
  env $v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
done
