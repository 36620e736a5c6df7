# The N = 8 / 4 / 2 / 1 runs of bench.py (default workload: batch + e2e + host fabric + mosaic + single image) on ONE
# 8-GPU box, plus what the box looks like from the host side.  usage: gpurun --gpus 8 -- bash tools/scale_box.sh [tag]
tag=${1:-s8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{ nvidia-smi topo -m; echo; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Hypervisor|Virtualization"; echo; (numactl -H 2>/dev/null || echo "numactl: not installed"); echo;
  for g in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | tr 'A-Z' 'a-z'); b=${b: -12}; echo "GPU$g $b numa_node=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) link=$(cat /sys/bus/pci/devices/$b/current_link_speed 2>/dev/null) x$(cat /sys/bus/pci/devices/$b/current_link_width 2>/dev/null)"; done; } > gpurun_out/${tag}_box.txt 2>&1
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --no-cpu > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
  echo "N=$N rc=$?"
done
timeout 400 python bench.py --gpus 1 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "N=1 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
tail -c 400 gpurun_out/${tag}_bench_n8.err; tail -c 300 gpurun_out/${tag}_bench_n8.json
