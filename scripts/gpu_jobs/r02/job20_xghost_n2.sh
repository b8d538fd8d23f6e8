# r02 job 20 (2 GPUs): buffer_x's exchange under the interior fine deposit -- multi-image parity (threads and NCCL), N=2 bench
set -x
timeout 600 python -m pytest tests/test_gpu_multi_image.py tests/test_gpu_nccl.py -m gpu -q -x > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02u_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-late --no-e2e > gpurun_out/r02u_bench_n2.log 2> gpurun_out/r02u_bench_n2.err; echo "rc=$?"
tail -3 gpurun_out/r02u_bench_n2.err
CUBE_GPU_SYNC_BUFFER_X=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-late --no-e2e > gpurun_out/r02u_bench_n2_sync.log 2> gpurun_out/r02u_bench_n2_sync.err; echo "rc=$?"
python - <<PY
import json
for f in ('r02u_bench_n2','r02u_bench_n2_sync'):
    d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], '%.4e'%d['value'])
PY
