python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f3.log 2>&1; tail -2 gpurun_out/pytest_f3.log
WPSGPU_POISON=1 python -m pytest tests -m gpu -x -q -k "metgrid or config or mirror or device" > gpurun_out/pytest_f3_poison.log 2>&1; tail -2 gpurun_out/pytest_f3_poison.log
python bench.py > gpurun_out/bench_f3.json 2> gpurun_out/bench_f3.err; tail -c 200 gpurun_out/bench_f3.json
WPSGPU_PRECISION=0 python bench.py --no-cpu --no-geogrid --no-e2e > gpurun_out/bench_f3_exact.json 2> gpurun_out/bench_f3_exact.err; python -c "import json; d=json.loads(open('gpurun_out/bench_f3_exact.json').read().strip().splitlines()[-1]); print('exact', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/geo_launches_f3.csv python tools/geo_only.py > /dev/null 2>&1
