python -m pytest tests -m gpu -x -q > gpurun_out/pytest_nmm.log 2>&1; tail -15 gpurun_out/pytest_nmm.log
