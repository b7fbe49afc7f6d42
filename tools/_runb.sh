timeout 600 python -m pytest tests/test_gpu_mpas.py -x -q > gpurun_out/pytest_mpas.log 2>&1; tail -30 gpurun_out/pytest_mpas.log
