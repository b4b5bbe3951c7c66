#!/bin/bash
# Gaussian-count / resolution sweep of one fwd+bwd frame (BASELINE.json configs[4] asks for it). Run on a GPU box.
out=${1:-gpurun_out/sweep.jsonl}
: > $out
for n in 250000 500000 1000000 2000000 3000000; do
  python tools/stage_times.py c3 presort 3 n_gauss=$n tight=1 2>/dev/null | tail -1 | sed "s/^{/{\"sweep\": \"n_gauss=$n 1920x1080 n=8\", /" >> $out
done
for wh in "1280 720" "2560 1440" "3840 2160"; do
  set -- $wh
  python tools/stage_times.py c3 presort 3 width=$1 height=$2 tight=1 2>/dev/null | tail -1 | sed "s/^{/{\"sweep\": \"n_gauss=1000000 $1x$2 n=8\", /" >> $out
done
for nv in 1 4 16; do
  python tools/stage_times.py c3 presort 3 n_virtual=$nv tight=1 2>/dev/null | tail -1 | sed "s/^{/{\"sweep\": \"n_gauss=1000000 1920x1080 n=$nv\", /" >> $out
done
python tools/stage_times.py c5 presort 2 n_frames=1 tight=1 2>/dev/null | tail -1 | sed "s/^{/{\"sweep\": \"c5 one frame: n_gauss=3000000 3840x2160 n=16\", /" >> $out
python tools/stage_times.py c2 presort 5 tight=1 2>/dev/null | tail -1 | sed "s/^{/{\"sweep\": \"c2: n_gauss=100000 800x800 n=4\", /" >> $out; cat $out
