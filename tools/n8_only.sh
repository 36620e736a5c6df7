timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --no-cpu > gpurun_out/s8b_bench_n8.json 2> gpurun_out/s8b_bench_n8.err; echo "N=8 rc=$?"
grep "bench rank 0" gpurun_out/s8b_bench_n8.err | tail -20; tail -c 300 gpurun_out/s8b_bench_n8.json
