# round 2: profiles.  ncu launch list of the default bench, full captures of the general, cells and exact-2x kernels, DRAM
# traffic of the dominant launch of configs 3, 4, 5, and compute-sanitizer runs.  Usage: bash tools/gpu_r2_prof.sh <tag>
TAG=${1:-r02p}
mkdir -p gpurun_out
B="--no-cpu --no-all-configs --no-verify --plugin-threads 0 --bands 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_config2.csv python bench.py --config 2 --steps 2 --warmup 1 $B > gpurun_out/${TAG}_ncu_launches.log 2>&1
bash tools/gpu_ncu.sh ${TAG}_ncu_up2x_config2 resample_up2x 2 -- --config 2 --steps 1 --warmup 1 $B > /dev/null
bash tools/gpu_ncu.sh ${TAG}_ncu_cells_config6 resample_cells 2 -- --config 6 --steps 1 --warmup 1 $B > /dev/null
bash tools/gpu_ncu.sh ${TAG}_ncu_strips_config8 resample_strips 1 -- --config 8 --steps 1 --warmup 1 $B > /dev/null
for c in 3 4 5; do
  K=resample_up2x; [ $c -eq 5 ] && K=resample_down
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$K -s 2 -c 1 --csv --log-file gpurun_out/${TAG}_traffic_config$c.csv python bench.py --config $c --steps 1 --warmup 1 $B > /dev/null 2>&1
  tail -4 gpurun_out/${TAG}_traffic_config$c.csv
done
# sanitizer: every kernel family once under memcheck (frame parity, row bands, device batch), racecheck on one case per family
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -q -x -p no:cacheprovider -k "(frame_matches and noise) or (row_bands and 3) or device_batch" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "frame_matches and noise and (c2_420 or c5_420 or up1p5_tap3_420p8 or down2to3_tap3_420p8 or irregular_up or up4to3_tap4)" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; tail -5 gpurun_out/${TAG}_sanitizer_racecheck.log
