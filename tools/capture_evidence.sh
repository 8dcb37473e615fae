#!/bin/bash
# Evidence capture on the GPU box (run under gpurun, ONE GPU):   bash tools/capture_evidence.sh r02 [streams=256] [step=93]
#   1. the bench command alone (must exit 0 before anything runs under ncu)
#   2. launch list of the same command: ncu --metrics gpu__time_duration.sum (cold-cache, serialised: compare SHARES)
#   3. ncu --set full of one steady-state step of every kernel class at the bench's stream count, one handle
#   4. single-stream lines (BASELINE.json configs 2 and 3): bench.py --streams 1 --handles 1
# Everything lands in gpurun_out/; tools/ncu_extract.py turns the report into profiles/<tag>_ncu_full.csv here.
TAG=${1:-r02}
S=${2:-256}
STEP=${3:-93}   # an odd steady-state step: lost-feature update AND prune update (the fleet prunes every other frame)
PART=${4:-all}  # list | full | all : the two ncu passes can run in separate gpurun calls (64 MiB limit on what comes back)
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --streams $S --handles 1 --steps 21 --warmup 3 --cpu-frames 1 --no-check"
FILTER='detect_kernel|klt_reg_kernel|pyr_down_bulk_kernel|pyr_down_strip_kernel|pyr_tail_kernel|be_select_kernel|be_gram_kernel|be_pchol_kernel|be_chol_kernel|be_gemm_kernel|be_feature_jac_kernel|be_feature_jac_prune_kernel|be_propagate_kernel|be_stack_kernel|be_scatter_kernel|be_triangulate_kernel|be_apply_kernel|fe_finish|fe_after_track|fe_after_stereo|fe_sieve'
if [ "$PART" != "full" ]; then
  $CMD > $OUT/${TAG}_plain_h1.json 2> $OUT/${TAG}_plain_h1.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain_h1.err; exit 1; }
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_all.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "launch list rc=$?"
  python tools/ncu_plan.py $OUT/${TAG}_launches_all.csv "$FILTER" $STEP > $OUT/${TAG}_plan.txt
  python tools/ncu_plan.py $OUT/${TAG}_launches_all.csv "$FILTER" $STEP --summary > $OUT/${TAG}_launch_summary.csv
  python bench.py --streams 1 --handles 1 --steps 40 --warmup 5 --cpu-frames 1 --no-check > $OUT/${TAG}_bench_1stream.json 2> $OUT/${TAG}_bench_1stream.err
  echo "1-stream rc=$?"
  gzip -f $OUT/${TAG}_launches_all.csv
fi
if [ "$PART" != "list" ]; then
  if [ -n "$5" ]; then SKIP=$5; COUNT=$6; else read SKIP COUNT < $OUT/${TAG}_plan.txt; fi
  echo "full capture: skip $SKIP count $COUNT"
  $CMD > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
  timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$FILTER" --launch-skip $SKIP --launch-count $COUNT -f -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
  echo "full capture rc=$?"
  # the summaries are extracted HERE (the report itself may be too large to travel back)
  python tools/ncu_extract.py $OUT/${TAG}_full.ncu-rep $S $OUT/${TAG} "ncu --set full --clock-control none --import-source on -k regex:<kernel classes> --launch-skip $SKIP --launch-count $COUNT; $CMD (ONE handle, $S streams = the bench's stream count; engine step $STEP: steady state, full windows, lost-feature update and prune update)"
  cp profiles/traffic.json $OUT/${TAG}_traffic.json
  cp profiles/inst.json $OUT/${TAG}_inst.json
  xz -T0 -2 -f $OUT/${TAG}_full.ncu-rep 2>/dev/null || gzip -f $OUT/${TAG}_full.ncu-rep
  SZ=$(du -sm $OUT | cut -f1)
  echo "gpurun_out: ${SZ} MiB"
  if [ "$SZ" -gt 60 ]; then rm -f $OUT/${TAG}_full.ncu-rep.*; echo "report dropped (too large to travel); summaries kept"; fi
fi
