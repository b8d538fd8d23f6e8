# r02 job 33: the CUDA path against the committed golden hashes (tests/golden/oracle_step_nc32.json)
set -x
python -m pytest tests/test_gpu_golden.py -m gpu -q -rs > gpurun_out/r02zj_pytest_golden.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02zj_pytest_golden.log
