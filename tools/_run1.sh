set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "binning or bin or sort or c2 or c3 or edge or stress or one_shot" 2>&1 | tail -8
echo "=== stage place"; timeout 300 python tools/stage_times.py c3 presort 5 2>&1 | tail -2
echo "=== stage cub"; CHS_BIN_VARIANT=1 timeout 300 python tools/stage_times.py c3 presort 5 2>&1 | tail -1
for ch in 2048 8192; do echo "=== chunk $ch"; CHS_BIN_CHUNK=$ch timeout 300 python tools/stage_times.py c3 presort 5 2>&1 | tail -1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'place_kernel|rects_kernel|column_scan|DeviceScan|emit_kernel|Onesweep|tile_offsets' -c 40 --csv --log-file gpurun_out/bin_launches.csv python tools/stage_times.py c3 presort 1 > gpurun_out/bin_ncu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo bench rc=$?
tail -c 1500 gpurun_out/bench_r1g.json
