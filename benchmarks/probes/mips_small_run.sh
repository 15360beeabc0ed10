set -x
timeout 300 python -m pytest tests/test_retrieval_gpu.py -x -q -m gpu -k "small_batch or topk or golden or ties or random or adversarial or invalid or async" > gpurun_out/small_tests.txt 2>&1
tail -15 gpurun_out/small_tests.txt
rm -f gpurun_out/small_probe.txt
for v in "GRB_MIPS_SMALL=1" "GRB_MIPS_SMALL_STRIDE=4" "GRB_MIPS_SMALL_STRIDE=16"; do
  env $v timeout 120 python benchmarks/probes/mips_small_probe.py >> gpurun_out/small_probe.txt 2>&1
done
cat gpurun_out/small_probe.txt
bash benchmarks/probes/mips_small_ncu.sh 2>/dev/null | tail -9
