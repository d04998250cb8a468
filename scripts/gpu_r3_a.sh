#!/bin/bash
# round-2 session 3: table rows in shared memory + L2 query prefetch in the TMA kernel: parity, then A/B timings
tag=${1:-t1}
out=gpurun_out/$tag
mkdir -p $out
{
timeout 120 python scripts/check_xattn.py cell_tma 2 768 224 8 7
timeout 120 python scripts/check_xattn.py cell_tma 1 1024 336 12 11
timeout 120 python scripts/check_xattn.py cell_tma 2 384 256 16 5
} > $out/check.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_tma.py tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -30 ) > $out/pytest.log
{
for lib in scripts/exp/libnaf_old.so "" scripts/exp/libnaf_qpf0.so scripts/exp/libnaf_qpf6.so scripts/exp/libnaf_pv2.so; do
  export NAF_B200_LIB=$lib
  [ -z "$lib" ] && unset NAF_B200_LIB
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 2
  timeout 120 python scripts/time_xattn.py 8 cell_tma 768 896 32 7 1
  timeout 120 python scripts/time_xattn.py 4 cell_tma 1024 1036 37 11 2
done
} > $out/time_xattn.log 2>&1
cat $out/check.log; tail -3 $out/pytest.log; cat $out/time_xattn.log
