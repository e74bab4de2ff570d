"""Summarise an ncu report of the flux kernel: key metrics + executed-instruction mix.
usage: python profiles/summarize.py raw.csv src.csv ncells_per_launch"""
import csv, collections, re, sys
raw, src, ncells = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__pcsamp_sample_count'] + \
       [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f'{k},{units[i]},{d[i]}')
rows = list(csv.reader(open(src)))
h = None
R = []
nsec = 0
for r in rows:
    if r and r[0] == 'Address':
        h = r
        nsec += 1
        continue
    if nsec == 1 and h and r and r[0].startswith('0x'):     # first kernel of the report only
        R.append(r)
ie, isrc, ismp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
tot = 0
byop, smp = collections.Counter(), collections.Counter()
for r in R:
    n = int(r[ie])
    tot += n
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
    op = m.group(2).split('.')[0] if m else '?'
    byop[op] += n
    smp[op] += int(r[ismp])
ncw = ncells / 32
print(f'# executed warp instructions per cell and stage: {tot / ncw:.1f} (static SASS instructions: {len(R)})')
print('opcode,per_cell_stage,percent,stall_samples')
for op, n in byop.most_common(28):
    print(f'{op},{n / ncw:.1f},{100 * n / tot:.1f},{smp[op]}')
