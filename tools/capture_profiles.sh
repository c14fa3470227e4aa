#!/bin/bash
# ncu captures behind profiles/ (run on the GPU box: gpurun -- 'bash tools/capture_profiles.sh'), then
# `python tools/make_profiles.py rNN` here turns gpurun_out/*.ncu-rep / launches.csv into the tracked summaries.
# Every profiled command is first run once WITHOUT ncu and must exit 0; numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "bench failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu1.log 2>&1
python tools/prof_one.py --w 3840 --h 2160 --spp 16 --reps 2 > gpurun_out/plain2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -f -o gpurun_out/prof_trace_4k16 \
    python tools/prof_one.py --w 3840 --h 2160 --spp 16 --reps 2 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -f -o gpurun_out/prof_trace_1080p1 \
    python tools/prof_one.py --spp 1 --reps 2 > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_(front|init|scene_box|morton|onesweep|rle|reorder|heap_up|nodes)" -c 10 -f -o gpurun_out/prof_build \
    python tools/prof_one.py --spp 1 --reps 1 > gpurun_out/ncu4.log 2>&1
python tools/prof_one.py --scene 10m --w 3840 --h 2160 --spp 16 --reps 2 > gpurun_out/plain5.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -f -o gpurun_out/prof_trace_4k16_10m \
    python tools/prof_one.py --scene 10m --w 3840 --h 2160 --spp 16 --reps 2 > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
