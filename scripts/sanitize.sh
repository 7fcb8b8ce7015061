#!/bin/bash
# Dev tool: compute-sanitizer passes over a small batch (run on the GPU box).
for tool in memcheck synccheck racecheck initcheck; do
  echo "=== $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/dump_subset.py 327 2947 5 6 7 8 9 10 11 12 13 14 2>&1 | grep -v "^$" | tail -15
done
