#!/bin/bash
# Development helper: one gpurun call = tests + smoke + bench (+ optional extras).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "not trsv" --maxfail=40 -p no:cacheprovider > gpurun_out/t_main.log 2>&1
echo "main tests exit $?" >> gpurun_out/status.txt
timeout 600 python -m pytest tests -m gpu -q -k "trsv" --maxfail=40 -p no:cacheprovider > gpurun_out/t_trsv.log 2>&1
echo "trsv tests exit $?" >> gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/status.txt
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/status.txt
"$@"
cat gpurun_out/status.txt
for f in gpurun_out/t_main.log gpurun_out/t_trsv.log gpurun_out/smoke.log; do tail -n 5 "$f"; done
