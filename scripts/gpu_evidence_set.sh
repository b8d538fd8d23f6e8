# The evidence set of a round in one gpurun call: the GPU test suite, bench.py and its
# reference arm at cfg 2, smoke(), the ncu launch list of the same bench command, one ncu --set full capture of the hot kernels at
# cfg 1 (reports come back in gpurun_out/).  Usage: gpurun --timeout 1500 -- "bash scripts/gpu_evidence_set.sh"
set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/ev_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/ev_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/ev_bench.log 2> gpurun_out/ev_bench.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/ev_bench.log; tail -3 gpurun_out/ev_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev_bench_ref.log 2> gpurun_out/ev_bench_ref.err; echo "ref rc=$?"; tail -c 1200 gpurun_out/ev_bench_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ev_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-late --profile > gpurun_out/ev_bench_ncu.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ev_traffic_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-late --profile > gpurun_out/ev_bench_ncu2.log 2>&1; echo "ncu traffic rc=$?"
KREGEX="k_fine_deposit|k_fft_|k_drift_|k_fine_kick|k_coarse_kick|k_coarse_cell|k_coarse_gather|k_mask_ext|k_flag"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KREGEX" -c 18 -f -o gpurun_out/ev_prof_full \
    python bench.py --nc 128 --nnt 2 --steps 1 --warmup 3 --no-cpu --no-e2e --no-late --profile > gpurun_out/ev_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/ | tail -5
