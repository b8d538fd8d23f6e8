# 2-GPU A/B of the coarse-mesh overlap with several images (NCCL), coarse stream at the highest priority
set -x
for v in A=1 CUBE_GPU_OVERLAP=1; do
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n2_$v.log 2> gpurun_out/bench_n2_$v.err; echo "bench2 $v rc=$?"
  python - <<PY
import json
l=json.loads(open("gpurun_out/bench_n2_$v.log").read().strip().splitlines()[-1])
print("$v", l["n_gpus"], round(l["ms_per_step"],2), "%.3e"%l["value"])
PY
done
for v in A=1 CUBE_GPU_COARSE_PRIO0=1; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err
  python - <<PY
import json
l=json.loads(open("gpurun_out/bench_$v.log").read().strip().splitlines()[-1])
print("N=1 $v", round(l["ms_per_step"],2))
PY
done
