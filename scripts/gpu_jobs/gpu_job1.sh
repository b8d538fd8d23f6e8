set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01a.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/pytest_gpu.log
