# r02 job 30 (2 GPUs): the bench exactly as the driver launches it for N = 2 (e2e with the streamed upload, late-time leg)
set -x
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02zg_bench_n2_default.log 2> gpurun_out/r02zg_bench_n2_default.err; echo "rc=$?"
tail -3 gpurun_out/r02zg_bench_n2_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02zg_bench_n2_default.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], '%.4e'%d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['serialised_breakdown_ms'], 'late', d['late_time']['ms_per_step'])
PY
