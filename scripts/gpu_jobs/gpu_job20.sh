# cells streamed out under particle_mesh: streamed-checkpoint test, ABI test, bench with the e2e leg
set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streamed or full_steps" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_e2e.log 2> gpurun_out/bench_e2e.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_e2e.log").read().strip().splitlines()[-1])
print(round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "%.4e"%l["e2e"]["value"])
PY
