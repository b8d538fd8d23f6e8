# r02 job 29: two different species (masses 0.9/0.1, hot second species) against the composed two-species oracle
set -x
python -m pytest tests/test_gpu_two_species.py -m gpu -q > gpurun_out/r02zf_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02zf_pytest.log
