#!/usr/bin/env bash
# Next GPU session: lanes per LP of the INIT launch (RS_INIT_SPREAD) on the 68,608-LP road grid, one GPU.
# spread 32 is issue-bound there (one lane per warp), spread 1 serialises the junctions.
#   gpurun --timeout 600 -- 'bash tools/_sweep_traffic_init.sh > gpurun_out/traffic_init_sweep.txt 2>&1'
cd "$(dirname "$0")/.."
for s in 32 16 8 4 2 1; do
	echo "== RS_INIT_SPREAD=$s"
	RS_INIT_SPREAD=$s timeout 150 python tools/multigpu_check_traffic.py traffic_grid_lp68608 3600 2>&1 | grep TRAFFIC_MULTIGPU
done
