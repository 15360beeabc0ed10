set -x
timeout 600 python -m pytest tests/test_retrieval_gpu.py tests/test_benchmark_shapes_gpu.py -x -q -m gpu -k "not c2_ and not c5_" > gpurun_out/sel_tests.txt 2>&1
tail -3 gpurun_out/sel_tests.txt
timeout 200 python benchmarks/kbench.py mipsc4 mipsshard8 2>&1 | tail -4
