# z=49 -> 0 at cfg 2 with the P(k) of the final state estimated on the GPU (1024^3 grid)
set -x
timeout 70 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 40 --pk > gpurun_out/evolve_cfg2_pk.jsonl 2> gpurun_out/evolve_cfg2_pk.err; echo "rc=$?"
tail -2 gpurun_out/evolve_cfg2_pk.jsonl | cut -c1-600; tail -3 gpurun_out/evolve_cfg2_pk.err
