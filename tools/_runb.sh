python -m pytest tests -m gpu -x -q -k "geogrid or geo or config3 or ref_read" > gpurun_out/pytest_geo.log 2>&1; tail -5 gpurun_out/pytest_geo.log
python tools/geo_only.py 2>&1 | tail -1 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/geo_launches.csv python tools/geo_only.py > /dev/null 2>&1
