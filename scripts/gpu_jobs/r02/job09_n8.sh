# r02 job 9 (8 GPUs): NCCL parity on 2/4/8 ranks, cfg 3 (8 x 512^3) bench, cfg 5 (8 x 1024^3 = 2048^3 particles) bench
set -x
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q > gpurun_out/r02j_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"
tail -4 gpurun_out/r02j_pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02j_bench_n8_cfg3.log 2> gpurun_out/r02j_bench_n8_cfg3.err; echo "cfg3 rc=$?"
tail -c 2500 gpurun_out/r02j_bench_n8_cfg3.log; tail -3 gpurun_out/r02j_bench_n8_cfg3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --nc 512 --nnt 8 --ic-tile 2 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r02j_bench_n8_cfg5.log 2> gpurun_out/r02j_bench_n8_cfg5.err; echo "cfg5 rc=$?"
tail -c 2500 gpurun_out/r02j_bench_n8_cfg5.log; tail -5 gpurun_out/r02j_bench_n8_cfg5.err
nvidia-smi --query-gpu=memory.used --format=csv | head -3
