for d in ${DBGS:-0 1 4 7}; do echo "dbg=$d"; GRB_FWD2_DBG=$d python benchmarks/kbench.py attnfwd 2>&1 | grep attn_fwd; done
