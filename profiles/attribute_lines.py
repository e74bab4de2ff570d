"""Executed instructions and stall samples of a kernel by SOURCE LINE: joins the SASS page of an ncu capture
(`ncu -i rep --page source --csv`, per-instruction counts) with `nvdisasm -g` of the same build (address -> file:line).
usage: python profiles/attribute_lines.py src.csv all.sass '<mangled kernel prefix>' [top N]
  all.sass:  cuobjdump -xelf all flux0_fast.o && nvdisasm -g flux_inst.sm_100a.cubin > all.sass"""
import collections, csv, os, re, sys
src_csv, sass, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(sass).read().split("\n")
start = next(n for n, l in enumerate(lines) if l.startswith(".text." + prefix))
addr2line, cur = {}, None
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h, R, nsec = None, [], 0
for r in rows:
    if r and r[0] == "Address":
        h = r; nsec += 1; continue
    if nsec == 1 and h and r and r[0].startswith("0x"):
        R.append(r)
ie, ismp = h.index("Instructions Executed"), h.index("# Samples")
base = int(R[0][0], 16)
agg = collections.defaultdict(lambda: [0, 0])
for r in R:
    loc = addr2line.get(int(r[0], 16) - base)
    agg[loc][0] += int(r[ie]); agg[loc][1] += int(r[ismp])
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"# {len(R)} SASS instructions, {ti} executed warp instructions, {ts} stall samples")
print("# percent of executed instructions, percent of stall samples, file:line, source")
cache = {}
for loc, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = loc if loc else ("?", 0)
    if f not in cache:
        cache[f] = open(f).read().split("\n") if os.path.exists(f) else []
    text = cache[f][l - 1].strip()[:100] if 0 < l <= len(cache[f]) else ""
    print(f"{100 * i / ti:5.1f}% {100 * s / ts:5.1f}% {os.path.basename(f)}:{l}  {text}")
