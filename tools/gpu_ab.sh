#!/bin/bash
# same box, back to back: older builds of the library vs the working tree (diagnostic)
for rep in 1 2; do
for lib in tools/_ab/libA_noclock.so tools/_ab/libB_clock.so plade_b200/libplade_b200.so; do
  echo "== $lib (rep $rep)"
  PLADE_AB_LIB=$lib timeout 300 python tools/concurrency_probe.py 2000000 1,4 5 2>&1 | grep "B="
done
done
