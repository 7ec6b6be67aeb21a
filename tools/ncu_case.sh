#!/bin/bash
# ncu --set full capture of one launch of each named case (see tools/prof_case.py).  The .ncu-rep embeds the whole 12 MB
# module (~50 MB per report), so it is reduced ON the GPU box to a one-line summary, the raw-metric CSV and the
# per-instruction source CSV; those land in gpurun_out/<tag>/.
# Usage: bash tools/ncu_case.sh <tag> <kernel-regex> case1 [case2 ...]
TAG=$1; KRE=$2; shift 2
O=gpurun_out/$TAG; mkdir -p $O
for c in "$@"; do
  python tools/prof_case.py $c 5 >> $O/timings.txt 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 2 -c 1 -f -o /tmp/$c python tools/prof_case.py $c 2 > $O/ncu_$c.log 2>&1
  python tools/ncu_summary.py /tmp/$c.ncu-rep > $O/$c.summary.txt 2>&1
  ncu -i /tmp/$c.ncu-rep --page raw --csv > $O/$c.raw.csv 2>/dev/null
  ncu -i /tmp/$c.ncu-rep --page source --csv > $O/$c.source.csv 2>/dev/null
  rm -f /tmp/$c.ncu-rep
done
cat $O/timings.txt; cat $O/*.summary.txt
