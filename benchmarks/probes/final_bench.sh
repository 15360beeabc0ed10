set -x
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err || { tail -20 gpurun_out/r2_bench_1gpu.err; exit 1; }
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['kernel'], d['roofline']['frac'])
r=d['retrieval']; print('C4', r['value'], r['roofline']['frac'], 'small', r['small_batch']['graph_ms'], r['small_batch']['graph_hbm_frac'], r['small_batch'].get('launches_per_call'), 'c3', r['c3']['ms'], r['c3']['graph_ms'], r['c3']['graph_hbm_frac'], r['c3'].get('launches_per_call'), r.get('clocks'))
ls=d['long_sequence']; print('C5 slice', ls['attention_slice']['fwd_frac'], ls['attention_slice']['bwd_frac'], 'step', ls['train_step']['ms_per_step'])
print('dropin', d['dropin_eager']['value'], 'loss_check', d['loss_check']['ok'], 'clocks', d['clocks'])
P
