# r02 job 26 (8 GPUs): cfg 5 (2048^3 particles) with the late-time leg: the run evolved to z=0 by the product's step loop, then timed
set -x
timeout 230 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --nc 512 --nnt 8 --ic-tile 2 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02zc_bench_n8_cfg5_late.log 2> gpurun_out/r02zc_bench_n8_cfg5_late.err; echo "rc=$?"
tail -4 gpurun_out/r02zc_bench_n8_cfg5_late.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zc_bench_n8_cfg5_late.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], '%.4e'%d['value'], d['late_time'])
PY
