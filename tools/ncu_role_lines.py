"""Stall samples of ONE warp role of the tensor-core frontend by (source line, opcode).
usage: ncu_role_lines.py <sass csv> <nvdisasm -g> <function substring> <lo> <hi> [top]   (byte-offset range from ncu_roles.py)"""
import csv, re, bisect, collections, sys
src_csv, disasm, func, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
top = int(sys.argv[6]) if len(sys.argv) > 6 else 25
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if r and r[0] == 'Address'); idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith('0x')]
base = min(int(r[0], 16) for r in data)
infunc = False; seq = []; last = None
for ln in open(disasm, errors='ignore'):
    if '.section' in ln:
        infunc = ('.text.' in ln and func in ln); continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: last = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m: seq.append((int(m.group(1), 16), last, m.group(2)))
offs = [a for a, _, _ in seq]
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(float(r[idx['# Samples']]) for r in data)
toti = sum(float(r[idx['Instructions Executed']]) for r in data)
agg = collections.Counter(); ins = collections.Counter(); st = collections.defaultdict(collections.Counter)
for r in data:
    off = int(r[0], 16) - base
    if not (lo <= off < hi): continue
    i = bisect.bisect_right(offs, off) - 1
    op = seq[i][2].split()
    key = (seq[i][1], op[1] if op[0].startswith('@') and len(op) > 1 else op[0])
    s = float(r[idx['# Samples']]); agg[key] += s; ins[key] += float(r[idx['Instructions Executed']])
    for c in stall_cols:
        try: st[key][c] += float(r[idx[c]])
        except Exception: pass
print("role share of all samples: %.1f%%, of all instructions %.1f%%" % (100 * sum(agg.values()) / tot, 100 * sum(ins.values()) / toti))
for k, v in agg.most_common(top):
    t3 = ", ".join("%s %.0f%%" % (c[6:], 100 * x / max(v, 1)) for c, x in st[k].most_common(3))
    print("%5.2f%% %-28s %-18s ins %5.2f%%  %s" % (100 * v / tot, "%s:%d" % k[0] if k[0] else '?', k[1], 100 * ins[k] / toti, t3))
