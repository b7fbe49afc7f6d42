python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f6.log 2>&1; tail -4 gpurun_out/pytest_f6.log
python bench.py > gpurun_out/bench_f6.json 2> gpurun_out/bench_f6.err; tail -2 gpurun_out/bench_f6.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_f6.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['geogrid']['ms_per_field'], d['clocks'])"
python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c5_f6.json 2> gpurun_out/bench_c5_f6.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_c5_f6.json').read().strip().splitlines()[-1]); print('c5', d['value'], d['ms_per_step'], d['roofline']['frac'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
