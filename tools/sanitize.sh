#!/bin/bash
# compute-sanitizer passes over the tcgen05 evaluation kernels and the Picard kernels (run on the GPU box through gpurun):
#   tools/sanitize.sh [outdir]
# memcheck on everything; racecheck and synccheck on the d = 20 and d = 100 resident-operand kernels and the K-streamed one.
# Summaries go to $OUT/sanitize_<tool>_<d>.log; profiles/r2_sanitizer.md quotes them.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  for d in 20 100 300; do
    log="$OUT/sanitize_${tool}_${d}.log"
    timeout 300 $CS --tool $tool --print-limit 20 \
        python tools/sanitize_target.py $d > "$log" 2>&1
    echo "== $tool d=$d rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' "$log" | tail -2 | tr '\n' ' ')"
  done
done
