#!/bin/bash
# One gpurun session: GPU tests, smoke, bench (both arms), kernel timings; optional ncu passes.
# Everything lands in gpurun_out/<tag>/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh r2a [ncu]'
tag=${1:-session}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $out/smi.txt 2>&1
nproc > $out/host.txt; free -g >> $out/host.txt
( timeout 900 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 ) > $out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -15 ) > $out/smoke.log
( timeout 600 python bench.py --e2e-d2h both 2> $out/bench_c2.err | tail -1 ) > $out/bench_c2.json
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2> $out/bench_ref.err | tail -1 ) > $out/bench_ref.json
{
  python scripts/time_xattn.py 8 auto 768 896 32 7 1
  python scripts/time_xattn.py 8 auto 768 896 32 7 2
  python scripts/time_xattn.py 1 auto 384 224 16 7 1
  python scripts/time_xattn.py 4 auto 1024 1036 37 11 1
  python scripts/time_xattn.py 4 auto 1024 1036 37 11 2
  python scripts/time_xattn.py 2 auto 768 1344 24 7 4
  python scripts/time_xattn.py 4 auto 768 2048 32 7 4
} > $out/time_xattn.log 2>&1
if [ "$2" = "ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs '' --e2e-d2h sample > $out/ncu_bench.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:xattn_cell_tcws -s 2 -c 1 -f -o $out/tcws_c2_rep2 \
     python scripts/ncu_xattn.py 8 cell_tcws 2 > $out/ncu_full.log 2>&1
fi
ls -la $out
