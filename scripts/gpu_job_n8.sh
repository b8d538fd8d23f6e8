# 8-GPU job: NCCL parity test on 2x2x2 images + bench at N=8 (one process per GPU)
set -x
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_gpu_nccl.py -x -q -k "8" 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"; tail -c 3500 gpurun_out/bench_n8.log; tail -8 gpurun_out/bench_n8.err
