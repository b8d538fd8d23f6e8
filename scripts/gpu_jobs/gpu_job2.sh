set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
