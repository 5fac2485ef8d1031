#!/bin/bash
# SASS evidence (no GPU needed): counts of the instruction families the design claims, per object file of the library.
# usage: scripts/sass_summary.sh > profiles/r02_sass_summary.txt
B=mri_interpolation_b200/build
printf "%-28s %8s %8s %8s %8s %8s %8s %10s %10s %10s %8s %8s\n" object UTCHMMA LDTM UTMALDG UTCBAR SYNCS HMMA REDG.F32x2 REDG.F32x4 LDGMC.ADD LDSM MUFU
for o in $B/*.o; do
  s=$(cuobjdump -sass $o 2>/dev/null)
  c() { echo "$s" | grep -c "$1"; }
  printf "%-28s %8d %8d %8d %8d %8d %8d %10d %10d %10d %8d %8d\n" $(basename $o .cu.o) $(c UTCHMMA) $(c "LDTM") $(c UTMALDG) $(c UTCBAR) $(c SYNCS) $(c "HMMA.16816") $(c "REDG.E.ADD.F32x2") $(c "REDG.E.ADD.F32x4") $(c "LDGMC.E.ADD") $(c LDSM) $(c MUFU)
done
echo
echo "multimem / red instructions in optim.cu.o:"
cuobjdump -sass $B/optim.cu.o | grep -oE "(LDGMC|STGMC|REDG|MULTIMEM)[A-Za-z0-9_.]*" | sort | uniq -c
echo
echo "kernel entry points in libmri_b200.so: $(cuobjdump -sass mri_interpolation_b200/libmri_b200.so | grep -c 'Function :')"
