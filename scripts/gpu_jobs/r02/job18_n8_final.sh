# r02 job 18 (8 GPUs): NCCL parity on 2/4/8 ranks, cfg 3 (8 x 512^3) with e2e, cfg 4 (two species, 8 images), cfg 5 (2048^3 particles)
set -x
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/r02s_topo.txt 2>&1
lscpu | egrep "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/r02s_lscpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q > gpurun_out/r02s_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"
tail -4 gpurun_out/r02s_pytest_nccl.log
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 "${@:4}" > gpurun_out/$3.log 2> gpurun_out/$3.err; echo "$3 rc=$?"; tail -c 1800 gpurun_out/$3.log; tail -3 gpurun_out/$3.err; }
run 600 29511 r02s_bench_n8_cfg3 --steps 5 --warmup 3 --no-cpu --no-late
run 300 29513 r02s_bench_n8_cfg4_species2 --species 2 --steps 5 --warmup 3 --no-cpu --no-late
run 900 29512 r02s_bench_n8_cfg5 --nc 512 --nnt 8 --ic-tile 2 --steps 3 --warmup 2 --no-cpu --no-e2e --no-late
nvidia-smi --query-gpu=memory.used --format=csv | head -3
