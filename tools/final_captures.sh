#!/bin/bash
# Round-end evidence of one build on one B200 (run under gpurun): bench lines, launch list, full ncu capture of the dominant launch,
# per-CTA timeline, protocol stress runs.  Everything lands in gpurun_out/ (copy what is to be kept into profiles/raw/).
set -u
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/bench_r2_final.json 2> $O/bench_r2_final.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r2_final_reference.json 2>> $O/bench_r2_final.err
python bench.py --config C2 --steps 20 --warmup 3 > $O/bench_r2_final_C2.json 2>> $O/bench_r2_final.err
python bench.py --config C4 --steps 20 --warmup 3 > $O/bench_r2_final_C4.json 2>> $O/bench_r2_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sample_|eval_|reduce_|point_weights|row_setup|rec_image" -c 400 --csv \
    --log-file $O/launches_r2_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:eval_tc_kernel<\(int\)2" -s 3 -c 1 -f \
    -o $O/prof_r2_eval_pde python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
python tools/tc_timeline.py > $O/tc_timeline_r2_final.txt 2>&1
{ echo "== d=100, 1000+200 centres, 40 point tiles per CTA, product library"; timeout 200 python tools/stress_eval.py 100 40 6 2>&1 | grep -v "^frame";
  echo "== d=20, 20+6 centres (one pair per class), 12 tiles per CTA"; timeout 200 python tools/stress_eval.py 20 12 4 -1 20 6 2>&1 | grep -v "^frame";
  echo "== d=100, 40+8 centres"; timeout 200 python tools/stress_eval.py 100 12 4 -1 40 8 2>&1 | grep -v "^frame";
  echo "== solver, d=20 n=rho=3, 1 200 test points, tcgen05 route beside the FP64 route"; timeout 300 python tools/stress_tc.py 20 3 1200 3 2>&1 | grep -v "^frame";
  echo "== solver, d=100 n=rho=4, 1 200 test points, 8 solves"; timeout 300 python tools/stress_tc.py 100 4 1200 8 2>&1 | grep -v "^frame"; } > $O/stress_r2_final.txt
tail -c 300 $O/bench_r2_final.err
ls -la $O/prof_r2_eval_pde.ncu-rep $O/launches_r2_final.csv
