# last sanity of the round: parity + streamed downloads + smoke with the final library
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_errors.py -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_last.log 2> gpurun_out/bench_last.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_last.log").read().strip().splitlines()[-1])
print(round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2))
PY
