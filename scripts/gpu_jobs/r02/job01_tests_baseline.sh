# r02 job 1: the whole GPU suite with the new nt=64 / all-plan / C-driver parity tests, then the round-1 baseline bench
set -x
nvidia-smi -L
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02a_pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench.log 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench.log; tail -3 gpurun_out/r02a_bench.err
