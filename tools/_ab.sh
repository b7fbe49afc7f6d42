export WPSGPU_LIB=$PWD/wps_b200/variants/libwpsgpu_evt.so
for v in 1 0 1 0; do
  echo "TIMING=$v"; WPSGPU_TIMING=$v python bench.py --steps 20 --warmup 3 --no-cpu --no-geogrid --no-e2e 2>gpurun_out/ab_t$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
done
