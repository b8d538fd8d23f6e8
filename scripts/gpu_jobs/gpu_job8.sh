# crowded-cell paths: parity tests, then the z=0 state of cfg 2 under several thresholds (sweep in evolve_bench.py)
set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 800 python scripts/evolve_bench.py --nc 256 --nnt 4 --max-seconds 560 --sweep "$1" > gpurun_out/evolve_cfg2.jsonl 2> gpurun_out/evolve_cfg2.err; echo "evolve2 rc=$?"
tail -20 gpurun_out/evolve_cfg2.jsonl | cut -c1-1500; tail -3 gpurun_out/evolve_cfg2.err
