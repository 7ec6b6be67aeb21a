#!/bin/bash
# One GPU-box visit: parity tests, bench, microbench, ncu launch list, ncu --set full of the hot kernels.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag> [stages...]   stages: test bench micro list full
TAG=${1:-run}; shift
STAGES=${*:-test bench micro list full}
O=gpurun_out/$TAG
mkdir -p $O
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt 2>&1
if has test; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
  tail -5 $O/pytest.log
fi
if has bench; then
  timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
  tail -c 3000 $O/bench.json
fi
if has micro; then
  timeout 600 python tools/microbench.py > $O/microbench.txt 2>&1; echo "micro exit $?"
  tail -40 $O/microbench.txt
fi
if has list; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
fi
if has full; then
  # one ncu --set full capture of the bench's hot kernels; reduced on the box (the .ncu-rep embeds the 12 MB module)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cache_write|cache_read|attn_mma' -s 6 -c 6 \
    -f -o /tmp/full python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1; echo "ncu full exit $?"
  python tools/ncu_summary.py /tmp/full.ncu-rep > $O/ncu_full_summary.txt 2>&1
  ncu -i /tmp/full.ncu-rep --page raw --csv > $O/ncu_full_raw.csv 2>/dev/null
  rm -f /tmp/full.ncu-rep
  cat $O/ncu_full_summary.txt
fi
