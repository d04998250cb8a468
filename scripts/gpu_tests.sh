#!/bin/bash
# quick gpurun session: GPU tests only (optionally a -k filter), logs under gpurun_out/<tag>/
tag=${1:-tests}; shift
out=gpurun_out/$tag
mkdir -p $out
( timeout 1200 python -m pytest tests -m gpu -q --durations=12 --tb=short "$@" 2>&1 | tail -400 ) > $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
