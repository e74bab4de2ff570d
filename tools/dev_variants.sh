#!/bin/bash
# development aid (GPU box): bench every libeb200_dev_*.so variant; usage: tools/dev_variants.sh "<size> <nb>" [variants...]
sz=${1:-"256 2"}; shift
set -- $sz "$@"; size=$1; nb=$2; shift 2
vars=${@:-$(ls gdtk_b200/csrc/libeb200_dev_*.so | sed 's/.*libeb200_dev_//; s/\.so//')}
for v in $vars; do
  export EB200_LIBRARY=$PWD/gdtk_b200/csrc/libeb200_dev_$v.so
  python bench.py --size $size --blocks-per-dim $nb --steps 10 --no-also --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v size $size waves=$EB200_CHUNK_WAVES min=$EB200_CHUNK_MIN:', round(d['value']/1e9,3), 'G; kernel ms', round(d['roofline']['kernel_ms_per_launch'],3))"
done
