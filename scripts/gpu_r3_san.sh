#!/bin/bash
out=gpurun_out/${1:-san3}
mkdir -p $out
( timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_tma.py 2>&1 | tail -25 ) > $out/compute_sanitizer_memcheck_tma.txt
( timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize_tma.py 2>&1 | tail -25 ) > $out/compute_sanitizer_racecheck_tma.txt
tail -12 $out/compute_sanitizer_memcheck_tma.txt; tail -12 $out/compute_sanitizer_racecheck_tma.txt
