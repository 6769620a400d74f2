#!/bin/bash
# usage (on the GPU box, via gpurun): bash tools_profile.sh <tag> [gib]
# Writes gpurun_out/<tag>_launches.csv (every launch with its device time) and full ncu captures of the two hot kernels.
TAG=${1:-r01}
G=${2:-8}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --gib $G --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hpcdc_scan -s 3 -c 1 -o gpurun_out/${TAG}_scan -f \
    python bench.py --gib $G --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_blake3_leaves -s 1 -c 1 -o gpurun_out/${TAG}_leaves -f \
    python bench.py --gib $G --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_leaves.log 2>&1
ls -la gpurun_out/
