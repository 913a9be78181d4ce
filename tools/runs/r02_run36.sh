#!/bin/bash
# cut thresholds of the sweep planner with the shared-memory ring kernel: PREALPS_BJ_SPLITA (a panel gets a whole CTA from a*q/8 k-blocks on)
# and PREALPS_BJ_CHUNKQ (a slice of a panel cut across CTAs = c*q k-blocks)
set -u
for cfg in "6 10" "4 10" "8 10" "6 6" "6 14" "4 6"; do
  set -- $cfg
  echo "== SPLITA=$1 CHUNKQ=$2"
  PREALPS_BJ_SPLITA=$1 PREALPS_BJ_CHUNKQ=$2 timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels " | cut -c1-60
  PREALPS_BJ_SPLITA=$1 PREALPS_BJ_CHUNKQ=$2 timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels " | cut -c1-60
done
