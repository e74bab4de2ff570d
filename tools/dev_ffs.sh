#!/bin/bash
# development aid (GPU box): bench the 2D forward-facing step with libeb200_dev_<variant>.so; usage: tools/dev_ffs.sh v1 v2 ...
for v in "$@"; do
  export EB200_LIBRARY=$PWD/gdtk_b200/csrc/libeb200_dev_$v.so
  python bench.py --workload ffs --steps 20 --no-also --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v ffs', round(d['value']/1e9,3), 'G; kernel ms', round(d['roofline']['kernel_ms_per_launch'],4), d['config']['library'][-60:])"
done
