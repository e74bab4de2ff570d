#!/bin/bash
# development aid (GPU box): parity growth + bench of the thermally-perfect-gas kernel with a DEV library; usage: tools/dev_tp.sh <variant>
v=${1:-tp}
export EB200_LIBRARY=$PWD/gdtk_b200/csrc/libeb200_dev_$v.so
timeout 300 python tools/dbg_growth.py tpg_box3d n=32 nb=2 steps=1,8 2>&1 | tail -2
timeout 300 python tools/dbg_growth.py tpg_box3d n=40 nb=1 steps=4 2>&1 | tail -1
timeout 300 python tools/dbg_growth.py tpg_box3d n=66 nb=2 steps=3 2>&1 | tail -1
timeout 300 python tools/dbg_growth.py tpg_ffs nx=120 ny=40 steps=20 2>&1 | tail -1
timeout 300 python bench.py --workload tpg --size 256 --blocks-per-dim 2 --steps 5 --no-also --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v tpg 256', d['value']/1e9, 'G; kernel ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'])"
