#!/bin/bash
# final check of the panel_window option (default build): bins independent of the window, auto on at 1M, event-timed kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 33 python scripts/check_panel_window.py 1000000 > gpurun_out/check49.log 2>&1; echo "exit=$?" >> gpurun_out/check49.log
tail -12 gpurun_out/check49.log | cut -c1-200
