#!/bin/bash
# Round-2 evidence run (one GPU): launch list, full ncu captures of the fused scan and the site pass,
# compute-sanitizer memcheck + racecheck of smoke().  Outputs under gpurun_out/.
set -u
OUT=gpurun_out
BENCH="python bench.py --motifs 12 --bp 1000000000 --steps 2 --warmup 3 --no-cpu --no-e2e --no-posteriors"
$BENCH > $OUT/r2_prof_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/r2_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_launches_config5.csv $BENCH > $OUT/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fast_scan_kernel -s 36 -c 12 -o $OUT/r2_fused_scan -f $BENCH > $OUT/r2_ncu_full.log 2>&1
echo "full scan rc=$?"
ncu --set full --clock-control none --import-source on -k regex:site_ -s 72 -c 6 -o $OUT/r2_site_pass -f $BENCH > $OUT/r2_ncu_site.log 2>&1
echo "full site rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/r2_memcheck.log 2>&1
echo "memcheck rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/r2_racecheck.log 2>&1
echo "racecheck rc=$?"
tail -3 $OUT/r2_memcheck.log $OUT/r2_racecheck.log
