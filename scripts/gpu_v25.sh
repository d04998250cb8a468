#!/bin/bash
out=gpurun_out/${1:-v25}
mkdir -p $out
( timeout 200 compute-sanitizer --tool memcheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_memcheck_union_bwdtc.txt
( timeout 200 compute-sanitizer --tool racecheck python scripts/sanitize_new.py 2>&1 | tail -25 ) > $out/compute_sanitizer_racecheck_union_bwdtc.txt
tail -4 $out/compute_sanitizer_memcheck_union_bwdtc.txt; tail -4 $out/compute_sanitizer_racecheck_union_bwdtc.txt
