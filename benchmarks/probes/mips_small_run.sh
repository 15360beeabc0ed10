set -x
timeout 600 python -m pytest tests/test_retrieval_gpu.py tests/test_benchmark_shapes_gpu.py -x -q -m gpu -k "not c2_ and not c5_" > gpurun_out/small_tests.txt 2>&1
tail -5 gpurun_out/small_tests.txt
rm -f gpurun_out/small_probe.txt
for v in "GRB_MIPS_SMALL=1" "GRB_MIPS_SMALL_STRIDE=4"; do
  env $v timeout 120 python benchmarks/probes/mips_small_probe.py >> gpurun_out/small_probe.txt 2>&1
done
cat gpurun_out/small_probe.txt
bash benchmarks/probes/mips_small_ncu.sh 2>/dev/null | tail -6
timeout 200 python benchmarks/kbench.py mipsc4 mipsshard8 2>&1 | tail -4
