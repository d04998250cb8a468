#!/bin/bash
out=gpurun_out/${1:-bwd}
mkdir -p $out
( timeout 600 python scripts/backward_sweep.py 2>&1 | tail -12 ) > $out/backward_sweep.log
cat $out/backward_sweep.log
