python -m pytest tests -m gpu -x -q -k "geogrid or config3 or golden or read_geogrid" > gpurun_out/pytest_ab.log 2>&1; tail -2 gpurun_out/pytest_ab.log
for v in base new base new; do
  if [ $v = base ]; then export WPSGPU_LIB=$PWD/wps_b200/variants/libwpsgpu_base.so; else unset WPSGPU_LIB; fi
  echo "VARIANT=$v"; python tools/geo_only.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_field'], d['topo_four_pt']['ms_per_field'])"
done
