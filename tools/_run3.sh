timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tight or binning or golden or c2" 2>&1 | tail -5
echo "=== stage square"; timeout 300 python tools/stage_times.py c3 presort 5 2>&1 | tail -2
echo "=== stage tight"; timeout 300 python tools/stage_times.py c3 presort 5 tight=1 2>&1 | tail -2
echo "=== stage tight + place"; CHS_BIN_VARIANT=2 CHS_BIN_CHUNK=2048 timeout 300 python tools/stage_times.py c3 presort 5 tight=1 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --bounds tight --no-cpu-baseline > gpurun_out/bench_tight.json 2> gpurun_out/bench_tight.err; echo bench rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_tight.json"))
print(d["value"], d["e2e"]["value"], d["e2e"]["serial_value"], d["roofline"]["frac"], d["config"]["isects_emitted_per_frame"], d["roofline"]["isects_per_launch"], d["extra"]["stage_ms_per_step"])
PY
