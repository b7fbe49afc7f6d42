python -m pytest tests/test_gpu_mpas.py -x -q > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-geogrid --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['fp32'], d['roofline']['frac'])"
