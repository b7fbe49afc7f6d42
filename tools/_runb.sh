python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f4.log 2>&1; tail -2 gpurun_out/pytest_f4.log
WPSGPU_GEO_BATCH=0 python -m pytest tests -m gpu -x -q -k "geogrid or config3 or golden" > gpurun_out/pytest_f4_nobatch.log 2>&1; tail -1 gpurun_out/pytest_f4_nobatch.log
python bench.py > gpurun_out/bench_f4.json 2> gpurun_out/bench_f4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_f4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['geogrid']['ms_per_field'], d['mpas_remap']['frac_of_hbm_peak'], d['clocks'])"
