#!/bin/bash
# A/B of the node kernel on the GPU box: rows in L2 (default) vs rows in shared memory
set -e
for B in 512 2048; do timeout 200 python tools/profile_run.py --batch $B --runs 3 | tail -2; done
MIQP_ROWS_SMEM=1 python planner-miqp_b200/build.py --force > /dev/null 2>&1
echo "--- rows in shared memory"
for B in 512 2048; do timeout 200 python tools/profile_run.py --batch $B --runs 3 | tail -2; done
python planner-miqp_b200/build.py --force > /dev/null 2>&1
