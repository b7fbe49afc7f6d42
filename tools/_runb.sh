python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f5.log 2>&1; tail -2 gpurun_out/pytest_f5.log
python bench.py > gpurun_out/bench_f5.json 2> gpurun_out/bench_f5.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_f5.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['geogrid']['ms_per_field'], d['geogrid']['value'], d['mpas_remap']['frac_of_hbm_peak'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/geo_launches_f5.csv python tools/geo_only.py > /dev/null 2>&1
