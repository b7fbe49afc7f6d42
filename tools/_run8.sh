N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload config5 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; tail -c 500 gpurun_out/bench_c5_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-geogrid > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; tail -c 300 gpurun_out/bench_c2_n$N.json
