python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; tail -2 gpurun_out/pytest_ab.log
for v in base new base new; do
  if [ $v = base ]; then export WPSGPU_LIB=$PWD/wps_b200/variants/libwpsgpu_base.so; else unset WPSGPU_LIB; fi
  echo "VARIANT=$v"; python bench.py --steps 20 --warmup 3 --no-cpu --no-geogrid --no-e2e 2>gpurun_out/ab_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['roofline']['breakdown_ms']; print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], b['k_metgrid_slowlist'], b['k_pack_win'])"
done
