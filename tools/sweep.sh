#!/bin/bash
# Tuning sweep of the wavefront kernel launch configuration on C2 (run under gpurun).
for blk in 256 128; do for sm in 1 0; do
  echo "== block=$blk scene_smem=$sm"
  RTB200_WF_BLOCK=$blk RTB200_WF_SCENE_SMEM=$sm python tools/render_once.py C2 3 2>&1 | tail -1
done; done
