#!/bin/bash
# scripts/gpu_profile.sh [tag] -- evidence run on the GPU box (one GPU): bench lines of the three extract workloads and both arms,
# the ncu launch list of the default bench command, and `ncu --set full` captures of every hot kernel for K1 / K2 / K4 and the
# matcher.  The captures are summarised ON THE BOX (raw metric pages as CSV, hottest source lines); only K1's and the matcher's
# .ncu-rep travel back (gpurun_out/ is limited to 64 MiB).  scripts/refresh_profiles.py <tag> turns the outputs into profiles/<tag>_*.
tag=${1:-r2}
O=gpurun_out
timeout 600 python bench.py > $O/bench_${tag}_k1.json 2> $O/bench_${tag}_k1.err; echo bench_k1_exit=$?
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${tag}_k1_20steps.json 2>/dev/null; echo bench_k1_20_exit=$?
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_${tag}_ref.json 2> $O/bench_${tag}_ref.err; echo ref_exit=$?
timeout 600 python bench.py --workload k2 --steps 200 --no-cpu-baseline --no-extras > $O/bench_${tag}_k2.json 2> $O/bench_${tag}_k2.err; echo bench_k2_exit=$?
timeout 600 python bench.py --workload k4 --steps 300 --no-cpu-baseline --no-extras > $O/bench_${tag}_k4.json 2> $O/bench_${tag}_k4.err; echo bench_k4_exit=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launches_${tag}.log 2>&1; echo launches_exit=$?
for wl in k1 k2 k4; do
  case $wl in k1) SC="-s 24 -c 12";; k2) SC="-s 22 -c 11";; k4) SC="-s 32 -c 16";; esac
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_fast_fused|k_blur|k_orient_desc|k_octree|k_resize" $SC -f \
      -o $O/prof_${tag}_${wl} python scripts/run_workload.py $wl 4 > $O/ncu_full_${tag}_${wl}.log 2>&1; echo full_${wl}_exit=$?
  ncu -i $O/prof_${tag}_${wl}.ncu-rep --page raw --csv > $O/ncu_raw_${tag}_${wl}.csv 2>/dev/null
  for k in k_fast_fused k_blur k_orient_desc k_octree k_resize; do
    echo "-- hottest source lines, $k" >> $O/ncu_hot_${tag}_${wl}.txt
    python scripts/ncu_source_hot.py $O/prof_${tag}_${wl}.ncu-rep $k 14 >> $O/ncu_hot_${tag}_${wl}.txt 2>/dev/null
  done
  [ $wl != k1 ] && rm -f $O/prof_${tag}_${wl}.ncu-rep
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_match_partial -s 5 -c 2 -f -o $O/prof_${tag}_matcher \
    python scripts/run_matcher.py 10 > $O/ncu_matcher_${tag}.log 2>&1; echo matcher_exit=$?
ncu -i $O/prof_${tag}_matcher.ncu-rep --page raw --csv > $O/ncu_raw_${tag}_matcher.csv 2>/dev/null
echo "-- hottest source lines, k_match_partial" > $O/ncu_hot_${tag}_matcher.txt
python scripts/ncu_source_hot.py $O/prof_${tag}_matcher.ncu-rep k_match_partial 14 >> $O/ncu_hot_${tag}_matcher.txt 2>/dev/null
du -sh $O; ls -la $O/*${tag}*
