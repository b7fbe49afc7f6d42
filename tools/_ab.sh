for v in base s10 s12 base s10 s12; do
  if [ $v = base ]; then unset WPSGPU_LIB; else export WPSGPU_LIB=$PWD/wps_b200/variants/libwpsgpu_$v.so; fi
  echo "VARIANT=$v"; python bench.py --steps 20 --warmup 3 --no-cpu --no-geogrid --no-e2e 2>gpurun_out/ab_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['roofline']['breakdown_ms']; print(d['ms_per_step'], d['roofline']['kernel_ms'], b['k_metgrid_slowlist'], b['k_pack_win'])"
done
