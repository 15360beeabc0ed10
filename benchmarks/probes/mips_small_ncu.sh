set -x
CMD="python benchmarks/probes/mips_small_probe.py c3"
$CMD > /dev/null 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/small_launches.csv $CMD > gpurun_out/small_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/small_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
names=[(r[ki][:70], r[vi]) for r in rows[1:]]
for n,v in names[-40:]: print(v, n)
P
