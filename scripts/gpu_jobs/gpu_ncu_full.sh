# one ncu --set full capture of the hot kernels on the cfg-1 workload (8 tiles of nt=64); the report comes back in gpurun_out/
set -x
KREGEX=${1:-"k_fine_deposit|k_fft_|k_drift_gather|k_fine_kick_p"}
COUNT=${2:-9}
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KREGEX" -c $COUNT -f -o gpurun_out/prof_full \
    python bench.py --nc 128 --nnt 2 --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/
