#!/bin/bash
# compute-sanitizer over the diagnostic groups at their small shapes: memcheck on every kernel group,
# racecheck + synccheck on the kernels with hand-rolled mbarrier / TMEM / grid-barrier protocols.  Summaries -> gpurun_out/sanitizer_*.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool group
  local out=gpurun_out/sanitizer_$1_$2.txt
  timeout ${3:-420} $CS --tool $1 --print-limit 20 --error-exitcode 9 python tools/gpu_diag.py $2 > $out.full 2>&1
  local rc=$?
  { echo "# compute-sanitizer --tool $1 python tools/gpu_diag.py $2   (exit code $rc; 9 = sanitizer errors, 124 = timeout)";
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race reported|hazard|barrier error|SUMMARY" $out.full | sort | uniq -c | head -20;
    echo "# diag result lines: $(grep -c '^PASS' $out.full) PASS, $(grep -c '^FAIL' $out.full) FAIL"; } > $out
  rm -f $out.full.keep; tail -c 3000 $out.full > $out.tail; rm -f $out.full
  cat $out
}
for g in ${GROUPS_MEM:-elementwise gemm_tn gemm_dw attention gcn0}; do run memcheck $g; done
for g in ${GROUPS_RACE:-gemm_tn gemm_dw attention gcn0}; do run racecheck $g 600; done
for g in ${GROUPS_SYNC:-gemm_tn gemm_dw attention gcn0}; do run synccheck $g; done
