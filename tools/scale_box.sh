set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/s8_bench_n$N.json 2> gpurun_out/s8_bench_n$N.err
  timeout 200 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --workload mosaic > gpurun_out/s8_mosaic_n$N.json 2> gpurun_out/s8_mosaic_n$N.err
  timeout 200 $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N --workload mosaic --fused > gpurun_out/s8_mosaic_fused_n$N.json 2> gpurun_out/s8_mosaic_fused_n$N.err
done
timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/s8_bench_n1.json 2> gpurun_out/s8_bench_n1.err
timeout 100 python bench.py --gpus 1 --workload mosaic --fused > gpurun_out/s8_mosaic_fused_n1.json 2> gpurun_out/s8_mosaic_fused_n1.err
tail -c 300 gpurun_out/s8_bench_n8.json; tail -c 300 gpurun_out/s8_mosaic_fused_n8.json
