#!/bin/bash
out=gpurun_out/${1:-probe}
mkdir -p $out
( timeout 300 python scripts/store_probe.py 2>&1 | tail -30 ) > $out/store_probe.log
cat $out/store_probe.log
