# r02 job 17: CUPTI timeline of two e2e steps (where the overlapped step loses time against its serialised parts)
set -x
python scripts/e2e_trace.py --out gpurun_out/r02r_e2e_trace.json > gpurun_out/r02r_e2e_trace.txt 2> gpurun_out/r02r_e2e_trace.err; echo "trace rc=$?"
tail -3 gpurun_out/r02r_e2e_trace.err
ls -la gpurun_out/r02r_e2e_trace.json
gzip -f gpurun_out/r02r_e2e_trace.json
python bench.py --steps 5 --warmup 3 --no-cpu --no-late --no-cfg1 --no-e2e > gpurun_out/r02r_bench.log 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02r_bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['phases_ms_per_step']['fine_deposit'])
PY
