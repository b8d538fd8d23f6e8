# z=49 -> z=0 evolution with per-window step timing (cfg 1 then cfg 2) + per-kernel DRAM traffic at cfg 2
set -x
timeout 400 python scripts/evolve_bench.py --nc 128 --nnt 2 --max-seconds 300 > gpurun_out/evolve_cfg1.jsonl 2> gpurun_out/evolve_cfg1.err; echo "evolve1 rc=$?"
tail -2 gpurun_out/evolve_cfg1.jsonl | cut -c1-600; tail -3 gpurun_out/evolve_cfg1.err
timeout 700 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 560 > gpurun_out/evolve_cfg2.jsonl 2> gpurun_out/evolve_cfg2.err; echo "evolve2 rc=$?"
tail -2 gpurun_out/evolve_cfg2.jsonl | cut -c1-600; tail -3 gpurun_out/evolve_cfg2.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/traffic_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --profile > gpurun_out/bench_ncu2.log 2>&1; echo "ncu rc=$?"
