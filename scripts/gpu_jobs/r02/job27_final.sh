# r02 job 27: final state of the round: whole GPU suite + the default bench line
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02zd_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02zd_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02zd_bench.log 2> gpurun_out/r02zd_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r02zd_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zd_bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], '%.4e'%d['value'], 'e2e', d['e2e']['ms_per_step'], 'late', d['late_time']['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
