# 2-GPU bench with the e2e leg (velocities streamed per tile batch, coarse mesh + NCCL on the priority stream)
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/bench_n2.log").read().strip().splitlines()[-1])
print(l["n_gpus"], round(l["ms_per_step"],2), "%.3e"%l["value"], "e2e", round(l["e2e"]["ms_per_step"],1), "%.3e"%l["e2e"]["value"])
PY
tail -3 gpurun_out/bench_n2.err
