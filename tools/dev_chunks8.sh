#!/bin/bash
# development aid (8-GPU box): the 512^3 bench under torchrun with different k-chunk policies (EB200_CHUNK_MIN / EB200_CHUNK_WAVES)
for cm in 32 16 22 43 64; do
  EB200_CHUNK_MIN=$cm python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$cm bench.py --gpus 8 --steps 20 --warmup 3 --no-also 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_min $cm:', round(d['value']/1e9,2), 'G; kernel ms', round(d['roofline']['kernel_ms_per_launch'],4), 'share', round(d['roofline']['kernel_share_of_step'],3))"
done
