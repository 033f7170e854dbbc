set -u
OUT=gpurun_out; TMP=/tmp/r2prof; mkdir -p $TMP
BENCH="python bench.py --motifs 12 --bp 1000000000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-posteriors"
$BENCH > $OUT/r2b_prof_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/r2b_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2b_launches_config5.csv $BENCH > $OUT/r2b_ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fast_scan_kernel -s 36 -c 12 -o $TMP/r2b_fused_scan -f $BENCH > $OUT/r2b_ncu_full.log 2>&1
echo "full scan rc=$?"
ncu -i $TMP/r2b_fused_scan.ncu-rep --page raw --csv > $OUT/r2b_ncu_fused_scan_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:site_ -s 72 -c 24 -o $TMP/r2b_site_pass -f $BENCH > $OUT/r2b_ncu_site.log 2>&1
echo "full site rc=$?"
ncu -i $TMP/r2b_site_pass.ncu-rep --page raw --csv > $OUT/r2b_ncu_site_pass_raw.csv 2>/dev/null
ncu -i $TMP/r2b_site_pass.ncu-rep --page source --csv --print-source sass > $TMP/site_sass.csv 2>/dev/null
python - <<'PY'
import csv
rd=csv.reader(open('/tmp/r2prof/site_sass.csv'))
blocks=[];cur=None
for r in rd:
    if r and r[0]=="Kernel Name":
        cur=[r]; blocks.append(cur); continue
    if cur is not None: cur.append(r)
# keep the first two site_score blocks only (size)
keep=[b for b in blocks if 'site_score' in b[0][1]][:2]
w=csv.writer(open('gpurun_out/r2b_ncu_site_score_sass.csv','w'))
for b in keep:
    for r in b: w.writerow(r[:12])
PY
du -sh $OUT
# the split kernel of the end-to-end leg (site records -> 6-byte form on the copy stream)
BENCH_E2E="python bench.py --motifs 12 --bp 1000000000 --steps 1 --warmup 1 --no-cpu --no-posteriors --e2e-steps 1"
ncu --set full --clock-control none -k regex:split_sites -c 4 -o $TMP/r2b_split -f $BENCH_E2E > $OUT/r2b_ncu_split.log 2>&1
echo "full split rc=$?"
ncu -i $TMP/r2b_split.ncu-rep --page raw --csv > $OUT/r2b_ncu_split_sites_raw.csv 2>/dev/null
