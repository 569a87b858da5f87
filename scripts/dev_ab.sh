#!/bin/bash
# A/B of the variant libraries built by scripts/dev_variants.py: C3 4096 children, two rounds, interleaved
mkdir -p gpurun_out/ab
for round in 1 2; do
  for f in bonmin_b200/_ab/libbnlp_*.so; do
    echo -n "$(basename $f) : "
    BNLP_LIB=$f python scripts/dev_prof.py 4096 2>&1 | tail -1
  done
done
