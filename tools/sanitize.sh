#!/bin/bash
# compute-sanitizer over one small launch of every kernel family (race / memory checking).
# Usage (on the GPU box):  bash tools/sanitize.sh [out_dir]      -> <out_dir>/sanitizer_<tool>.txt (+ a summary)
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
for TOOL in memcheck synccheck racecheck; do
  for CASE in ${SANITIZE_CASES:-sim sscd vit fid pair}; do
    LOG="$OUT/sanitizer_${TOOL}_${CASE}.txt"
    timeout ${SANITIZE_TIMEOUT:-300} $CS --tool $TOOL --print-limit 20 --launch-timeout 120 \
        python tools/sanitize_case.py $CASE > "$LOG" 2>&1
    echo "rc=$?" >> "$LOG"
  done
done
{
  echo "compute-sanitizer summary ($(date -u +%FT%TZ), $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1))"
  for f in "$OUT"/sanitizer_*_*.txt; do
    echo "== $(basename $f): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $f | tr '\n' ' ') $(grep -c 'sanitize_case done' $f) run(s) completed, $(tail -1 $f)"
  done
} > "$OUT/sanitizer_summary.txt"
cat "$OUT/sanitizer_summary.txt"
