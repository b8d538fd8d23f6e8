# tests + bench + ncu launch list + one ncu --set full capture; everything lands in gpurun_out/
set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
KREGEX=${1:-"k_fine_deposit|k_fft_|k_drift_|k_fine_kick_p|k_coarse_kick_p|k_coarse_deposit|k_f2max_rows"}
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KREGEX" -c 14 -f -o gpurun_out/prof_full \
    python bench.py --nc 128 --nnt 2 --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/
