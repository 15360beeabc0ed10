# full ncu captures of the small-batch plan's PRIVATE pass and select kernel on the C3 shape.
# Launch order of mips_scores_sm100_kernel in the probe: [0] GMAX, [1] PRIVATE of the graph's warm-up on
# an all-zero query batch (degenerate: every item a hit), [2] GMAX, [3] PRIVATE of the first real call.
set -x
CMD="python benchmarks/probes/mips_small_probe.py c3"
$CMD > /dev/null 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mips_scores_sm100_kernel -s 3 -c 1 -f -o gpurun_out/r2_mips_small_private $CMD > gpurun_out/ncu_small_private.log 2>&1
[ -n "$SKIP_SELECT" ] || timeout 300 ncu --set full --clock-control none --import-source on -k regex:mips_small_select -s 1 -c 1 -f -o gpurun_out/r2_mips_small_select $CMD > gpurun_out/ncu_small_select.log 2>&1
ls -la gpurun_out/*.ncu-rep
