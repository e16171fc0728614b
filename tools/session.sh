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
    trace1) for p in "f32 f64" "f64 f64" "f32 f32"; do set -- $p
              TRSV_VARIANT=1 ACCBLAS_LIB=$PWD/accessor-blas_b200/libaccblas_b200_dev.so timeout 120 python tools/trsv_trace.py 16384 $1 $2 > gpurun_out/${tag}_trace1_$1_$2.log 2>&1; echo "trace1 $p rc=$?"; grep "single-CTA" gpurun_out/${tag}_trace1_$1_$2.log; done ;;
    pytest) timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log ;;
    pytestall) timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest.log ;;
    sweep)  timeout 600 python tools/sweep_dot_fill.py > gpurun_out/${tag}_sweep.log 2>&1; echo "sweep rc=$?" ;;
    multi)  timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/${tag}_multi.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/${tag}_multi.log ;;
    multibench) (nvidia-smi topo -m; ls /sys/devices/system/node; cat /sys/bus/pci/devices/*/numa_node | sort | uniq -c; nproc; numactl -H) > gpurun_out/${tag}_topology.txt 2>&1
            for n in 1 2 4 8; do
              if [ $n -le $(nvidia-smi -L | wc -l) ]; then
                if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --steps 50 --warmup 5 --no-detail > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
                else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 50 --warmup 5 --no-detail > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; fi
                echo "bench n=$n rc=$?"; fi
            done ;;
    trsvab) timeout 600 python tools/trsv_ab.py > gpurun_out/${tag}_trsv_ab.log 2>&1; echo "trsvab rc=$?"; tail -30 gpurun_out/${tag}_trsv_ab.log ;;
    pytrsv) timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -x -k trsv > gpurun_out/${tag}_pytrsv.log 2>&1; echo "pytrsv rc=$?"; tail -5 gpurun_out/${tag}_pytrsv.log ;;
    sweeppool) timeout 600 python tools/sweep_dot_pool.py > gpurun_out/${tag}_sweeppool.log 2>&1; echo "sweeppool rc=$?"; cat gpurun_out/${tag}_sweeppool.log ;;
    sweeppool26) DOT_LOG2N=26 timeout 600 python tools/sweep_dot_pool.py > gpurun_out/${tag}_sweeppool26.log 2>&1; echo "sweeppool26 rc=$?"; cat gpurun_out/${tag}_sweeppool26.log ;;
    pydot) timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -x -k "dot" > gpurun_out/${tag}_pydot.log 2>&1; echo "pydot rc=$?"; tail -5 gpurun_out/${tag}_pydot.log ;;
    sizes)  timeout 900 python tools/size_sweep.py > gpurun_out/${tag}_sizes.log 2>&1; echo "sizes rc=$?"; tail -3 gpurun_out/${tag}_sizes.log ;;
    bench)  timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" ;;
    benchref) timeout 600 python bench.py --impl reference --steps 10 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "benchref rc=$?" ;;
    smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    compare) timeout 900 python tools/compare_reference.py > gpurun_out/${tag}_compare.log 2>&1; echo "compare rc=$?" ;;
    ncutrsv) ACCBLAS_PROFILE_TRSV=f64:f32,f64:f64 timeout 300 python tools/profile_target.py trsv > gpurun_out/${tag}_ncu_plain.log 2>&1 && \
             ACCBLAS_PROFILE_TRSV=f64:f32,f64:f64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trsv_cluster_kernel -s 1 -c 1 -o gpurun_out/${tag}_trsv_f32 python tools/profile_target.py trsv > gpurun_out/${tag}_ncu.log 2>&1; echo "ncutrsv rc=$?"
             ACCBLAS_PROFILE_TRSV=f64:f64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trsv_cluster_kernel -s 1 -c 1 -o gpurun_out/${tag}_trsv_f64 python tools/profile_target.py trsv >> gpurun_out/${tag}_ncu.log 2>&1; echo "ncutrsv2 rc=$?" ;;
    ncutrsv1) ACCBLAS_PROFILE_TRSV=f64:f32 timeout 300 python tools/profile_target.py trsv > gpurun_out/${tag}_ncu1_plain.log 2>&1 && \
             ACCBLAS_PROFILE_TRSV=f64:f32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trsv_kernel -s 1 -c 1 -o gpurun_out/${tag}_trsv1 python tools/profile_target.py trsv > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncutrsv1 rc=$?"
             ncu -i gpurun_out/${tag}_trsv1.ncu-rep --page source --csv > gpurun_out/${tag}_trsv1_source.csv 2>/dev/null
             ncu -i gpurun_out/${tag}_trsv1.ncu-rep --page raw --csv > gpurun_out/${tag}_trsv1_raw.csv 2>/dev/null ;;
    profile) export ACCBLAS_PROFILE_TRSV=f64:f32,f64:f64,f32:f32,f64:f16
             timeout 600 python tools/profile_target.py > gpurun_out/${tag}_prof_plain.log 2>&1 && \
             timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemv_stream|dot_stream|trsv_kernel|trsv_cluster|fill_linear|convert_kernel" -o gpurun_out/${tag}_prof python tools/profile_target.py > gpurun_out/${tag}_prof_ncu.log 2>&1; echo "profile rc=$?"
             ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_raw.csv 2>/dev/null; rm -f gpurun_out/${tag}_prof.ncu-rep ;;
    launches) timeout 600 python bench.py --steps 3 --warmup 3 --no-detail --no-config5 > gpurun_out/${tag}_launch_plain.log 2>&1 && \
             timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-detail --no-config5 > gpurun_out/${tag}_launch_ncu.log 2>&1; echo "launches rc=$?" ;;
    *) echo "unknown step $step" ;;
  esac
done
