# 2-GPU job: NCCL parity test + bench at N=1 and N=2
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_nccl.py -x -q 2>&1 | tail -30
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; tail -c 1500 gpurun_out/bench_n1.log; tail -3 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; tail -c 3000 gpurun_out/bench_n2.log; tail -15 gpurun_out/bench_n2.err
