python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f2.log 2>&1; tail -3 gpurun_out/pytest_f2.log
python bench.py > gpurun_out/bench_f2.json 2> gpurun_out/bench_f2.err; tail -c 300 gpurun_out/bench_f2.json
python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c5_f2.json 2> gpurun_out/bench_c5_f2.err; tail -c 400 gpurun_out/bench_c5_f2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_f2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-geogrid --no-e2e > gpurun_out/ncu_l_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:strip16 -s 3 -c 1 -o gpurun_out/r02g_strip16 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-geogrid --no-e2e > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/r02g_strip16.ncu-rep
