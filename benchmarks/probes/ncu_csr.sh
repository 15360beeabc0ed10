set -x
CMD="python bench.py --steps 2 --warmup 3 --skip-retrieval --skip-cpu-baseline --skip-long-sequence"
$CMD > /dev/null 2>&1 || exit 1
for k in csr_rows_kernel csr_cols_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 3 -c 1 -f -o gpurun_out/r2_$k $CMD > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/r2_csr_*
