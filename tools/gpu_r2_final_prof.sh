# round 2, final code: plan build times, ncu launch list of the default bench, full captures of the exact-2x kernel (config 2),
# the strips-only launch of config 2, DRAM traffic of the dominant launch of configs 2-5, compute-sanitizer on the planned
# strips and the bulk-copy staging.  Usage: bash tools/gpu_r2_final_prof.sh <tag>
TAG=${1:-r02z}
mkdir -p gpurun_out
B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
for c in 2 2 4 3 5; do JINCRESIZE_B200_PLAN_TIMING=1 timeout 300 python bench.py --config $c --steps 1 --warmup 1 $B 2>&1 >/dev/null | grep "strip plan"; done > gpurun_out/${TAG}_plan_timing.txt; cat gpurun_out/${TAG}_plan_timing.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config2.csv python bench.py --config 2 --steps 2 --warmup 1 $B > gpurun_out/${TAG}_ncu_launches.log 2>&1
bash tools/gpu_ncu.sh ${TAG}_ncu_up2x_config2 resample_up2x 2 -- --config 2 --steps 1 --warmup 1 $B > /dev/null
bash tools/gpu_ncu.sh ${TAG}_ncu_up2x_config2_strips resample_up2x 2 -- --config 2 --steps 1 --warmup 1 $B --parts 2 > /dev/null
for c in 2 3 4 5; do
  K=resample_up2x; [ $c -eq 5 ] && K=resample_down
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$K -s 2 -c 1 --csv --log-file gpurun_out/${TAG}_traffic_config$c.csv python bench.py --config $c --steps 1 --warmup 1 $B > /dev/null 2>&1
  tail -3 gpurun_out/${TAG}_traffic_config$c.csv | cut -c1-200
done
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider -k "(frame_matches and noise) or (row_bands and 3) or device_batch or strip_plan or bulk_copy" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider -k "(frame_matches and noise and (c2_420 or c5_420 or up1p5_tap3_420p8 or down2to3_tap3_420p8 or down2to3_tap4_y16 or irregular_up or up4to3_tap4)) or (strip_plan and (c2_420 or c4_rgbps or up4to3 or c5_420 or up1p5_tap3_420p8)) or bulk_copy" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.log
