O=gpurun_out; tag=r2b; wl=k1
timeout 170 ncu --set full --import-source on --clock-control none -k regex:"k_fast_fused|k_blur|k_orient_desc|k_octree|k_resize" -s 24 -c 12 -f \
    -o $O/prof_${tag}_${wl} python scripts/run_workload.py $wl 4 > $O/ncu_full_${tag}_${wl}.log 2>&1; echo full_${wl}_exit=$?
ncu -i $O/prof_${tag}_${wl}.ncu-rep --page raw --csv > $O/ncu_raw_${tag}_${wl}.csv 2>/dev/null
for k in k_fast_fused k_blur k_orient_desc k_octree k_resize; do
  echo "-- hottest source lines, $k" >> $O/ncu_hot_${tag}_${wl}.txt
  python scripts/ncu_source_hot.py $O/prof_${tag}_${wl}.ncu-rep $k 14 >> $O/ncu_hot_${tag}_${wl}.txt 2>/dev/null
done
rm -f $O/prof_${tag}_${wl}.ncu-rep
wc -l $O/ncu_raw_${tag}_${wl}.csv
