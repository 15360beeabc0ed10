# 2-GPU sanity of the data-parallel paths after the step fusions: the multi-GPU tests, then the bench line
set -x
timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu > gpurun_out/mg_tests.txt 2>&1
tail -3 gpurun_out/mg_tests.txt | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err || tail -20 gpurun_out/r2_bench_2gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
r=d['retrieval']; print('C4', r['value'], r['ms_per_step'], r['e2e']['value'])
print('C5', d['long_sequence']['train_step']['ms_per_step'] if d.get('long_sequence') else None)
P
