# r02 job 25 (2 GPUs): the late-time leg of the bench with several images (small images: nc=128, nnt=2)
set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --nc 128 --nnt 2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02zb_bench_n2_late.log 2> gpurun_out/r02zb_bench_n2_late.err; echo "rc=$?"
tail -4 gpurun_out/r02zb_bench_n2_late.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zb_bench_n2_late.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['late_time'])
PY
