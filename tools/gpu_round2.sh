#!/bin/bash
# Round-2 evidence run: parity tests, bench, ncu launch list of the bench, ncu --set full of the tcgen05 attention kernel
# (plain and interpolating) and of the fused write / read kernels, microbench.  Usage: bash tools/gpu_round2.sh <tag> [stages]
TAG=${1:-r02}; shift
STAGES=${*:-test bench list full micro}
O=gpurun_out/$TAG
mkdir -p $O
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt 2>&1
if has test; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
  tail -4 $O/pytest.log
fi
if has bench; then
  timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err; echo "reference exit $?"
fi
if has list; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-triton > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
fi
if has full; then
  for c in attn attni; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 1 -f -o /tmp/$c python tools/prof_case.py $c 2 > $O/ncu_$c.log 2>&1
    python tools/ncu_summary.py /tmp/$c.ncu-rep > $O/$c.summary.txt 2>&1
    ncu -i /tmp/$c.ncu-rep --page source --csv > $O/$c.source.csv 2>/dev/null
    python tools/ncu_roles.py $O/$c.source.csv > $O/${c}_roles.txt 2>&1
    gzip -9f $O/$c.source.csv; rm -f /tmp/$c.ncu-rep
  done
  timeout 900 ncu --set full --clock-control none -k regex:'cache_write3|cache_read2' -s 6 -c 4 -f -o /tmp/wr python bench.py --steps 3 --warmup 3 --no-cpu --no-attention > $O/ncu_wr.log 2>&1
  python tools/ncu_summary.py /tmp/wr.ncu-rep > $O/write_read.summary.txt 2>&1; rm -f /tmp/wr.ncu-rep
  cat $O/*.summary.txt
fi
if has micro; then
  timeout 600 python tools/microbench.py > $O/microbench.txt 2>&1; echo "micro exit $?"
fi
