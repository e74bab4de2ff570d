#!/bin/bash
# development aid (GPU box): parity growth + bench with a DEV library; usage: tools/dev_check.sh [variant] [quick]
v=${1:-}; [ -n "$v" ] && v=_$v
export EB200_LIBRARY=$PWD/gdtk_b200/csrc/libeb200_dev$v.so
python tools/dbg_growth.py box3d n=32 nb=2 steps=1,16 2>&1 | tail -2
python tools/dbg_growth.py box3d n=40 nb=1 steps=4,12 2>&1 | tail -2
if [ -z "$2" ]; then
python tools/dbg_growth.py box3d n=66 nb=2 steps=3 2>&1 | tail -1
python tools/dbg_growth.py box3d n=72 nb=1 steps=3 gasdynamic_update_scheme=tvd-rk3 2>&1 | tail -1
python tools/dbg_growth.py ffs nx=60 ny=20 steps=20 2>&1 | tail -1
python tools/dbg_growth.py ffs nx=300 ny=100 steps=10 2>&1 | tail -1
fi
for sz in "256 2" "512 4"; do set -- $sz
python bench.py --size $1 --blocks-per-dim $2 --steps 10 --no-also --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v size $1', d['value']/1e9, 'G; kernel ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'])"
done
