#!/bin/bash
# tiny panels (<= 32 steps) riding in the regular sweep launch instead of their own launch when a level has fewer than PREALPS_BJ_TINYFOLD of them
set -u
for f in 512 2500 6000 20000 1000000; do
  echo "== TINYFOLD=$f"
  PREALPS_BJ_TINYFOLD=$f timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels " | cut -c1-60
  PREALPS_BJ_TINYFOLD=$f timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels " | cut -c1-60
done
