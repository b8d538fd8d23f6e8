# r02 job 12: CUBEnu order + vmax(3), asynchronous buffer_v (in-process images), whole suite
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02m_pytest.log
