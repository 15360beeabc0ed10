set -x
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_retrieval_gpu.py tests/test_pipeline_gpu.py tests/test_benchmark_shapes_gpu.py tests/test_input_path_gpu.py -x -q -m gpu > gpurun_out/fusion_tests.txt 2>&1
tail -4 gpurun_out/fusion_tests.txt | cut -c1-300
grep -n "Error\|assert\|FAILED" gpurun_out/fusion_tests.txt | head -20 | cut -c1-300
timeout 300 python bench.py --steps 300 --warmup 5 --skip-retrieval --skip-cpu-baseline --skip-long-sequence > gpurun_out/fusion_bench.json 2> gpurun_out/fusion_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/fusion_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('final_loss'), d.get('loss_check'))
P
