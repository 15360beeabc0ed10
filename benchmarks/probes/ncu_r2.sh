# r2 ncu captures (one GPU).  Each program is run plainly first; the profiler pass follows only if it exits 0.
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export KBENCH_FAST=1
run() {  # name, kernel regex, skip count, command...
  local name=$1 regex=$2 skip=$3; shift 3
  "$@" > /dev/null 2>&1 || { echo "plain run of $name failed"; return; }
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f -o gpurun_out/r2_$name "$@" > gpurun_out/ncu_$name.log 2>&1
}
run attn_short_fwd hstu_attn_short_fwd 2 python benchmarks/kbench.py attnc2
run attn_short_bwd hstu_attn_short_bwd 1 python benchmarks/kbench.py attnc2
run attn_fwd2_long hstu_attn_fwd2 1 python benchmarks/kbench.py attnfwd
run attn_bwd_long hstu_attn_bwd_sm100 0 python benchmarks/kbench.py attnbwd
run proj_gemm proj_gemm_kernel 10 python bench.py --steps 2 --warmup 3 --skip-retrieval --skip-cpu-baseline --skip-long-sequence
run ssl_bwd ssl_bwd_vec 3 python bench.py --steps 2 --warmup 3 --skip-retrieval --skip-cpu-baseline --skip-long-sequence
run mips_c4_last mips_scores_sm100 4 python benchmarks/kbench.py mipsc4
python bench.py --steps 2 --warmup 3 --skip-retrieval --skip-cpu-baseline --skip-long-sequence > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench_train.csv python bench.py --steps 2 --warmup 3 --skip-retrieval --skip-cpu-baseline --skip-long-sequence > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
