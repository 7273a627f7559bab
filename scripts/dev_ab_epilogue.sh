#!/bin/bash
# developer A/B: kernel configuration + per-op time of the last decoder block's final convolution, fp16 vs fp16c plan
for prec in fp16 fp16c; do
  echo "==== $prec"
  PSSR_V3_VERBOSE=1 python scripts/dev_time_net.py 64 $prec 2>&1 | grep -E "forward|op4[01] |v3: 128x128 n=64 kb=1[124]"
done
