"""Which mbarrier waits of the tensor-core frontend carry the stall samples (ncu SASS page + nvdisasm -g), per call site."""
import csv, re, sys, bisect, collections
src_csv, disasm, func = sys.argv[1], sys.argv[2], sys.argv[3]
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
src = open('/root/repo/speech-intent-recognizer_b200/csrc/frontend_tc.cu').read().splitlines()
_pw = next(i for i, l in enumerate(src, 1) if 'void pipe_wait(uint64_t* bar' in l)
PW = (_pw, next(i for i in range(_pw, _pw + 40) if src[i - 1].startswith('}')))
# a wait = samples on instructions inside pipe_wait / mbar_* helper lines; attribute to the NEXT frontend_tc.cu line >= 230 that follows in address order
def is_wait(loc, sass):
    if not loc: return False
    return (loc[0] == 'tc_common.cuh' and loc[1] < 130) or (loc[0] == 'frontend_tc.cu' and PW[0] <= loc[1] <= PW[1])
tot = sum(float(r[idx['# Samples']]) for r in data)
agg = collections.Counter(); insn = collections.Counter()
for r in data:
    off = int(r[0], 16) - base
    i = bisect.bisect_right(offs, off) - 1
    if i < 0 or not is_wait(seq[i][1], seq[i][2]): continue
    j = i
    while j >= 0 and not (seq[j][1] and seq[j][1][0] == 'frontend_tc.cu' and seq[j][1][1] >= 230): j -= 1
    k = i
    while k < len(seq) and not (seq[k][1] and seq[k][1][0] == 'frontend_tc.cu' and seq[k][1][1] >= 230): k += 1
    key = (seq[j][1][1] if j >= 0 else None, seq[k][1][1] if k < len(seq) else None)
    agg[key] += float(r[idx['# Samples']]); insn[key] += float(r[idx['Instructions Executed']])
ti = sum(float(r[idx['Instructions Executed']]) for r in data)
print("wait samples %.1f%% of all, wait instructions %.1f%% of all" % (100 * sum(agg.values()) / tot, 100 * sum(insn.values()) / ti))
for (a, b), v in agg.most_common(14):
    print("%5.2f%% smp %5.2f%% ins  between line %s and %s | %s" % (100 * v / tot, 100 * insn[(a, b)] / ti, a, b, src[b - 1].strip()[:100] if b else ''))
