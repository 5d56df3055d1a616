#!/bin/bash
# development aid: time / check several builds of libpinnstep.so (tools/bin/lib_<tag>.so) in one GPU call
#   tools/ab_variants.sh "<tags for timing>" "<tags for accuracy>" "<tag for pytest>"
out=gpurun_out/ab.log
: > $out
for t in $1; do
  lib=tools/bin/lib_$t.so; [ "$t" = base ] && lib=pinns_fluid_dynamics_b200/lib/libpinnstep.so
  echo "== $t" >> $out
  PINN_LIBPINNSTEP=$lib python tools/quick_time.py cavity_steady 1000000 >> $out 2>&1
  PINN_LIBPINNSTEP=$lib python tools/quick_time.py poiseuille_flow 10000 >> $out 2>&1
done
for t in $2; do
  lib=tools/bin/lib_$t.so; [ "$t" = base ] && lib=pinns_fluid_dynamics_b200/lib/libpinnstep.so
  echo "== accuracy $t" >> $out
  PINN_LIBPINNSTEP=$lib python tools/accuracy_check.py cavity_steady 20000 2>&1 | tail -1 >> $out
  PINN_LIBPINNSTEP=$lib python tools/accuracy_check.py poiseuille_flow 10000 2>&1 | tail -1 >> $out
done
if [ -n "$3" ]; then
  echo "== pytest $3" >> $out
  PINN_LIBPINNSTEP=tools/bin/lib_$3.so python -m pytest tests -m gpu -x -q -k "not tensor_core_engine and not layered and not bfgs and not limited_memory" 2>&1 | tail -5 >> $out
fi
cat $out
