# round-end evidence: GPU suite, full bench (value, e2e, cpu baseline), reference arm, ncu launch list, per-kernel DRAM traffic
# at cfg 2, ncu --set full of the hot kernels at cfg 1, smoke()
set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 1200 gpurun_out/bench_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_ncu2.log 2>&1; echo "ncu traffic rc=$?"
KREGEX="k_fine_deposit|k_fft_|k_drift_|k_fine_kick|k_coarse_kick|k_coarse_deposit|k_mask_ext"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KREGEX" -c 16 -f -o gpurun_out/prof_full \
    python bench.py --nc 128 --nnt 2 --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/ | tail -5
