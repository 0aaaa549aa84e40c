#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 16" "16 8" "8 8" "8 16"; do set -- $cfg; echo "== timeline FWD=$1 BWD=$2"; CCB_BATCH_FWD=$1 CCB_BATCH_BWD=$2 timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd:|bwd:"; done
