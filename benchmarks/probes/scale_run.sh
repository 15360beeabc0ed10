# bench.py at N GPUs (argument), both arms at N=1; outputs under gpurun_out/
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_1gpu.json 2> gpurun_out/r2_bench_ref_1gpu.err
  python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 benchmarks/probes/dp_timeline.py 60 2>&1 | grep -v -i "warn\|triggered\|return Variable\|^$\|\*\*\*\|OMP" | tail -$((N+2)) > gpurun_out/r2_dp_timeline_${N}gpu.txt
fi
tail -c 600 gpurun_out/r2_bench_${N}gpu.json
