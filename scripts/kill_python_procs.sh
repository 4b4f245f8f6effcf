#!/bin/bash
# Kill the training processes that hold the GPUs (by PID taken from nvidia-smi, never by name pattern).
for pid in $(nvidia-smi --query-compute-apps=pid --format=csv,noheader 2>/dev/null | sort -u); do
    echo "killing $pid ($(ps -o comm= -p "$pid" 2>/dev/null))"
    kill "$pid" 2>/dev/null
done
