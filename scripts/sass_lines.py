"""Map ncu per-SASS-instruction stall samples to source lines (nvdisasm -g line info).
    python scripts/sass_lines.py <ncu sass csv> <nvdisasm -g -c output> <kernel substring>"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
blk = [r for r in rows[2:] if len(r) >= len(hdr)]
cur = fn = None
dis = []
for line in open(sys.argv[2]):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m and fn and sys.argv[3] in fn:
        dis.append((cur, m.group(2).strip()))
n = min(len(blk), len(dis))
print("ncu instrs", len(blk), "nvdisasm instrs", len(dis))
agg, tot, exe = {}, 0, {}
for i in range(n):
    s = int(blk[i][isamp] or 0)
    tot += s
    agg[dis[i][0]] = agg.get(dis[i][0], 0) + s
    exe[dis[i][0]] = exe.get(dis[i][0], 0) + int(blk[i][iex] or 0)
byfile = {}
for (f, l), s in agg.items():
    byfile[f] = byfile.get(f, 0) + s
print("samples", tot, {k: f"{100*v/tot:.1f}%" for k, v in byfile.items()})
for (f, l), s in sorted(agg.items(), key=lambda kv: -kv[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    print(f"{f:16s} {l:4d} {s:7d} {100*s/tot:5.1f}%  exec={exe[(f,l)]}")
