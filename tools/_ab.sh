python -m pytest tests/test_gpu_geogrid.py -x -q > gpurun_out/pytest_ab.log 2>&1; tail -12 gpurun_out/pytest_ab.log
