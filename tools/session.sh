#!/bin/bash
# One gpurun session: each step under its own timeout, logs under gpurun_out/.
# usage: tools/session.sh <tag> step [step ...]
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
for step in "$@"; do
  case $step in
    trsvq)  timeout 300 python tools/trsv_check.py quick > gpurun_out/${tag}_trsv_check.log 2>&1; echo "trsv_check rc=$?" ;;
    trsv)   timeout 900 python tools/trsv_check.py > gpurun_out/${tag}_trsv_check.log 2>&1; echo "trsv_check rc=$?" ;;
    trace)  for p in "f32 f64" "f64 f64" "f32 f32"; do set -- $p
              ACCBLAS_LIB=$PWD/accessor-blas_b200/libaccblas_b200_dev.so timeout 120 python tools/trsv_trace.py 16384 $1 $2 > gpurun_out/${tag}_trace_$1_$2.log 2>&1; echo "trace $p rc=$?"; done ;;
    pytest) timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log ;;
    pytestall) timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest.log ;;
    sweep)  timeout 600 python tools/sweep_dot_fill.py > gpurun_out/${tag}_sweep.log 2>&1; echo "sweep rc=$?" ;;
    bench)  timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" ;;
    benchref) timeout 600 python bench.py --impl reference --steps 10 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "benchref rc=$?" ;;
    smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    compare) timeout 900 python tools/compare_reference.py > gpurun_out/${tag}_compare.log 2>&1; echo "compare rc=$?" ;;
    *) echo "unknown step $step" ;;
  esac
done
