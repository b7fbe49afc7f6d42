python tools/mpas_only.py > gpurun_out/mpas_only.json 2> gpurun_out/mpas_only.err; cat gpurun_out/mpas_only.json; tail -3 gpurun_out/mpas_only.err
ncu --set full --clock-control none --import-source on -k regex:k_mpas_remap_planes -s 2 -c 1 -o gpurun_out/r02_mpas_planes -f python tools/mpas_only.py > gpurun_out/ncu_mpas.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_mpas --csv --log-file gpurun_out/mpas_launches.csv python tools/mpas_only.py > /dev/null 2>&1
