# r02 job 21 (8 GPUs): cfg 3 with buffer_x's exchange under the interior fine deposit, against the synchronous exchange
set -x
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-late --no-e2e > gpurun_out/$2.log 2> gpurun_out/$2.err; echo "$2 rc=$?"; }
run 29541 r02v_bench_n8_async_x
CUBE_GPU_SYNC_BUFFER_X=1 run 29542 r02v_bench_n8_sync_x
python - <<PY
import json
for f in ('r02v_bench_n8_async_x','r02v_bench_n8_sync_x'):
    d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], '%.4e'%d['value'])
PY
