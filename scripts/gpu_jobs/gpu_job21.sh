# velocities streamed per tile batch under particle_mesh (cube_gpu_stream_vp): tests + bench e2e with and without
set -x
python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "streamed or abi or drift_then" 2>&1 | tail -4
for v in "" "--no-stream-vp"; do
  python bench.py --steps 5 --warmup 3 --no-cpu $v > gpurun_out/bench_e2e$v.log 2> gpurun_out/bench_e2e$v.err; echo "bench rc=$?"
  python - <<PY
import json
l=json.loads(open("gpurun_out/bench_e2e$v.log").read().strip().splitlines()[-1])
print("$v", round(l["ms_per_step"],2), "e2e", round(l["e2e"]["ms_per_step"],2), "%.4e"%l["e2e"]["value"])
PY
done
tail -3 gpurun_out/bench_e2e.err
