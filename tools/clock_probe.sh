#!/bin/bash
# Sample clocks / power while the hot path runs for ~10 s (dev tool).
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm --format=csv
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu --format=csv,noheader -lms 250 > gpurun_out/clocks_probe.csv &
SMI=$!
python tools/run_once.py RealESRGAN_x4plus 720 1280 2 100 > gpurun_out/clock_run.log 2>&1
kill $SMI
sort gpurun_out/clocks_probe.csv | uniq -c | sort -rn | head -15
