# cfg-5-sized image on one GPU (1024^3 particles, nc=512, nnt=8: 512 tiles): does the HBM layout hold, what does a step cost
set -x
free -g | head -2
timeout 900 python bench.py --nc 512 --nnt 8 --ic-tile 2 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_cfg5.log 2> gpurun_out/bench_cfg5.err; echo "bench cfg5 rc=$?"
tail -c 3000 gpurun_out/bench_cfg5.log; tail -5 gpurun_out/bench_cfg5.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
