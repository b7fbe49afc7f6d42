python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f1.log 2>&1; tail -3 gpurun_out/pytest_f1.log
python bench.py > gpurun_out/bench_f1.json 2> gpurun_out/bench_f1.err; tail -c 600 gpurun_out/bench_f1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_f1.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-geogrid --no-e2e > gpurun_out/ncu_l_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:strip16 -s 3 -c 1 -o gpurun_out/r02f_strip16 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-geogrid --no-e2e > gpurun_out/ncu_f1.log 2>&1
ls -la gpurun_out/r02f_strip16.ncu-rep
