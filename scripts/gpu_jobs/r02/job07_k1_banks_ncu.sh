# r02 job 7: stash bank fix of the key+chain pass; ncu --set full of the particle kernels at cfg 1
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_tile.py tests/test_gpu_multi_image.py -m gpu -q -x > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02g_pytest.log
for v in A=1; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02g_bench_$v.log 2> gpurun_out/r02g_bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02g_bench_$v.log').read().strip().splitlines()[-1])
    print('$v', d['ms_per_step'], {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})
except Exception as e:
    print('$v failed', e); print(open('gpurun_out/r02g_bench_$v.err').read()[-600:])
PY
done
KREGEX="k_fine_deposit|k_drift_|k_fine_kick|k_coarse_kick|k_coarse_cell|k_flag"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KREGEX" -c 12 -f -o gpurun_out/r02g_prof \
    python bench.py --nc 128 --nnt 2 --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/r02g_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02g_prof* 
