#!/bin/bash
# one GPU call that produces the round's evidence from the committed build (copied into profiles/ afterwards by hand / make_profiles.py)
set -x
tag=${1:-r02}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_${tag}_final.log 2>&1; tail -3 $o/pytest_${tag}_final.log
python bench.py --steps 20 --warmup 5 > $o/bench_${tag}_final.log 2> $o/bench_${tag}_final.err; tail -c 600 $o/bench_${tag}_final.log
python tools/quick_time.py cavity_steady 1000000 > $o/qt_${tag}_final.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_tc_kernel -s 3 -c 1 -o $o/prof_${tag}_final -f python tools/quick_time.py cavity_steady 1000000 > $o/ncu_${tag}_final.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $o/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_${tag}_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $o/ncu_launch.log 2>&1
PINN_LIBPINNSTEP=tools/bin/libpinnstep_prof.so python tools/tc_phase_profile.py 1000000 > $o/phase_${tag}_final.log 2>&1
PINN_LIBPINNSTEP=tools/bin/libpinnstep_prof.so python tools/tc_trace.py 1000000 3 > $o/trace_${tag}_final.log 2>&1
cat $o/qt_${tag}_final.log $o/phase_${tag}_final.log
