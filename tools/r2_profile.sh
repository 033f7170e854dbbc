#!/bin/bash
# Round-2 evidence run (one GPU): launch list, full ncu captures of the fused scan and the site pass,
# compute-sanitizer memcheck + racecheck of smoke().  Outputs under gpurun_out/ (CSV pages; the
# .ncu-rep files are kept only while they fit gpurun's 64 MiB return limit).
set -u
OUT=gpurun_out
TMP=/tmp/r2prof; mkdir -p $TMP
BENCH="python bench.py --motifs 12 --bp 1000000000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-posteriors"
if [ "${SKIP_NCU:-0}" != "1" ]; then
$BENCH > $OUT/r2_prof_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/r2_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_launches_config5.csv $BENCH > $OUT/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fast_scan_kernel -s 36 -c 12 -o $TMP/r2_fused_scan -f $BENCH > $OUT/r2_ncu_full.log 2>&1
echo "full scan rc=$?"
ncu -i $TMP/r2_fused_scan.ncu-rep --page raw --csv > $OUT/r2_ncu_fused_scan_raw.csv 2>/dev/null
ncu -i $TMP/r2_fused_scan.ncu-rep --page source --csv --print-source sass > $OUT/r2_ncu_fused_scan_sass.csv 2>/dev/null
ncu --set full --clock-control none -k regex:site_ -s 72 -c 6 -o $TMP/r2_site_pass -f $BENCH > $OUT/r2_ncu_site.log 2>&1
echo "full site rc=$?"
ncu -i $TMP/r2_site_pass.ncu-rep --page raw --csv > $OUT/r2_ncu_site_pass_raw.csv 2>/dev/null
ls -la $TMP
fi
SMOKE='import __graft_entry__ as g; g.smoke(); print("smoke ok")'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "$SMOKE" > $OUT/r2_memcheck.log 2>&1
echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "$SMOKE" > $OUT/r2_racecheck.log 2>&1
echo "racecheck rc=$?"
tail -n 4 $OUT/r2_memcheck.log; tail -n 4 $OUT/r2_racecheck.log
du -sh $OUT
