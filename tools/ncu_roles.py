"""Share of stall samples / executed instructions per warp role of the tensor-core frontend (by SASS address range)."""
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
marks = {}
for i, l in enumerate(src, 1):
    if 'if (warp < 4) {' in l and 'A' not in marks: marks['A'] = i
    if 'else if (warp == kWarpMma)' in l: marks['MMA'] = i
    if 'else if (warp < kWarpD0)' in l: marks['C'] = i
    if 'else if (warp < kWarpF0)' in l: marks['DE'] = i
    if 'F: tile -> global' in l: marks['F'] = i
    if 'else if (warp == kWarpPub)' in l: marks['PUB'] = i
order = sorted(marks.items(), key=lambda kv: kv[1])
def first_addr(lo, hi):
    c = [a for a, l, _ in seq if l and l[0] == 'frontend_tc.cu' and lo <= l[1] < hi]
    return min(c) if c else None
bounds = []
for k, (name, line) in enumerate(order):
    hi = order[k + 1][1] if k + 1 < len(order) else 10 ** 6
    # first address of a line inside the role body that is NOT shared setup: take the max of the role's minimal addresses after the previous bound
    bounds.append((first_addr(line + 1, min(line + 12, hi)), name))
bounds = sorted(b for b in bounds if b[0] is not None)
def role(off):
    r = 'setup'
    for a, k in bounds:
        if off >= a: r = k
    return r
def is_wait(loc):
    return loc and ((loc[0] == 'tc_common.cuh' and loc[1] < 130) or (loc[0] == 'frontend_tc.cu' and PW[0] <= loc[1] <= PW[1]))
S = collections.Counter(); W = collections.Counter(); I = collections.Counter(); WI = collections.Counter()
for r in data:
    off = int(r[0], 16) - base
    i = bisect.bisect_right(offs, off) - 1
    ro = role(off); s = float(r[idx['# Samples']]); n = float(r[idx['Instructions Executed']])
    S[ro] += s; I[ro] += n
    if is_wait(seq[i][1]): W[ro] += s; WI[ro] += n
ts = sum(S.values()); ti = sum(I.values())
print(bounds)
nw = {'A': 4, 'MMA': 1, 'C': 4, 'DE': 4, 'F': 3}
for k in ('setup', 'A', 'MMA', 'C', 'DE', 'F'):
    if k in S:
        share = S[k] / ts
        busy = (S[k] - W[k]) / S[k] if S[k] else 0
        print(f"{k:6s} samples {share*100:5.1f}% (waiting {W[k]/ts*100:5.1f}% -> busy {busy*100:4.0f}% of its time)  instr {I[k]/ti*100:5.1f}% (wait loops {WI[k]/ti*100:5.1f}%)")
