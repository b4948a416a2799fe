#!/bin/bash
mkdir -p gpurun_out
for n in 1000 1500 1750 2000 2500; do echo "== nodes $n"; timeout 300 python scripts/k1_lab.py --nodes $n --graphs $((512000/n)) --variants lean,blocks:4,blocks:28 2>&1 | tail -3; done
