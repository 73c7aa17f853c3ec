#!/bin/bash
# Evidence for "hand-written Blackwell code": per-kernel histogram of the tensor-core / TMEM / TMA opcodes in the built library and the
# resource usage of every kernel.  Runs in the build container (no GPU): bash tools/sass_summary.sh
set -e
LIB=comfyui-float_optimized_b200/csrc/libfmt_b200.so
OUT1=profiles/r02_sass_histogram.txt
OUT2=profiles/r02_ptxas_resources.txt
{
  echo "# cuobjdump -sass $LIB | per-kernel counts of tcgen05 / TMEM / TMA / bulk-copy / cluster opcodes (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), sm_100a)"
  echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load, UTMAREDG = TMA tensor reduce-add,"
  echo "# UTMASTG = TMA tensor store, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, CCTL.IVALL = L1 invalidate behind ld.acquire"
  cuobjdump -sass $LIB | awk '
    /Function :/ { fn=$3 }
    { for (i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|LDTM|UTCBAR|UTMALDG|UTMAREDG|UTMASTG|UBLKCP|SYNCS|USETMAXREG|UTCATOMSWS|CCTL|UCGABAR_ARV|REDG|ATOMG)/) { op=$i; sub(/;$/,"",op); c[fn" "op]++ } }
    END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{n=$NF; $NF=""; printf "%6d  %s\n", n, $0}' | sed 's/(fmt::[A-Za-z]*Params)//'
} > $OUT1
{
  echo "# cuobjdump -res-usage $LIB  (registers / stack / static shared memory per kernel; dynamic shared memory is set at launch)"
  cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/ Function \([^:]*\):/\1/' | awk '{print $0}' | c++filt | sed 's/Fatbin elf code://'
} > $OUT2
wc -l $OUT1 $OUT2
