# r02 job 31: the bench state drifts away from the ICs as steps accumulate: 5 against 25 timed steps
set -x
for k in 5 25; do python bench.py --steps $k --warmup 3 --no-cpu --no-late --no-cfg1 --no-e2e > gpurun_out/r02zh_bench_steps$k.log 2> gpurun_out/r02zh_bench_steps$k.err; echo "rc=$?"; done
python - <<PY
import json
for k in (5,25):
    d=json.loads(open('gpurun_out/r02zh_bench_steps%d.log'%k).read().strip().splitlines()[-1])
    print(k, d['ms_per_step'], {a: round(b,2) for a,b in d['phases_ms_per_step'].items() if a.startswith('drift')})
PY
