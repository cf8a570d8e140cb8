#!/bin/bash
# round 2, GPU session O: the cluster team kernel against the cooperative team kernel (bit identity, us per iteration)
mkdir -p gpurun_out
L=mcrat_b200/csrc/libmcrat_b200.so
cp $L /tmp/libnocluster.so
# A = cooperative team (MCRAT_B200_NO_CLUSTER), B = cluster team; ab_compare passes the environment through, so two calls
( for cfg in "C2 100000 16 3000" "C1 10000 1 3000" "C5 100000 16 3000" "C3 100000 16 1000" "C2 100000 64 2000" "C5 1000000 16 1000"; do
  set -- $cfg
  echo "== $cfg"
  MCRAT_B200_NO_CLUSTER=1 timeout 200 python tools/ab_compare.py $L:persistent $L:streamed $1 $2 $3 $4 2>&1 | tail -3 | head -1 | sed "s/^/cooperative /"
  cp /tmp/ab_0.npy /tmp/ab_coop.npy
  timeout 200 python tools/ab_compare.py $L:persistent $L:streamed $1 $2 $3 $4 2>&1 | tail -3 | sed "s/^/cluster /"
  python -c "
import numpy as np
a=np.load('/tmp/ab_coop.npy'); b=np.load('/tmp/ab_0.npy')
print('cooperative vs cluster:', 'BIT-IDENTICAL' if a.tobytes()==b.tobytes() else 'DIFFERENT')"
done ) 2>&1 | tee gpurun_out/ab_o.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4
