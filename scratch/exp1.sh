#!/bin/bash
for e in "" "64,16,5" "64,16,6" "64,16,7" ""; do
  if [ -z "$e" ]; then unset NUTS_B200_ENGINE; else export NUTS_B200_ENGINE=$e; fi
  timeout 300 python scratch/bench_quick.py 2>&1 | tail -1
done
