# 8-GPU job: bench at N=8 (one process per GPU), coarse mesh overlapped on the high-priority stream; NCCL parity test after it
set -x
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"; tail -c 1500 gpurun_out/bench_n8.log; tail -4 gpurun_out/bench_n8.err
timeout 300 python -m pytest tests/test_gpu_nccl.py -x -q -k "8" 2>&1 | tail -3
