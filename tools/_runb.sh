ncu --set full --clock-control none --import-source on -k regex:"k_pack_win|k_metgrid_slowlist" -s 6 -c 2 -o gpurun_out/r02_packwin_slow -f python bench.py --steps 1 --warmup 3 --no-cpu --no-geogrid --no-e2e > gpurun_out/ncu_ps.log 2>&1
ls -la gpurun_out/r02_packwin_slow.ncu-rep
