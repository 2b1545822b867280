#!/bin/bash
# scripts/gpu_profile.sh [tag] -- evidence run on the GPU box: bench (both arms), ncu launch list, one `ncu --set full`
# capture per kernel.  Outputs under gpurun_out/; scripts/refresh_profiles.py turns them into profiles/.
tag=${1:-r1}
python bench.py > gpurun_out/bench_${tag}_k1.json 2> gpurun_out/bench_${tag}_k1.err; echo bench_exit=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo ref_exit=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1; echo launches_exit=$?
ncu --set full --import-source on --clock-control none -k regex:"k_fast_fused|k_blur|k_orient_desc|k_octree" -s 8 -c 8 -f -o gpurun_out/prof_${tag}_main \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1; echo full_exit=$?
ncu --set full --import-source on --clock-control none -k regex:"k_resize" -s 14 -c 7 -f -o gpurun_out/prof_${tag}_resize \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2_${tag}.log 2>&1; echo full2_exit=$?
ls -la gpurun_out/*${tag}*
