"""Aggregate an ncu SASS source page (csv) per CUDA source line using nvdisasm -g line markers.
usage: ncu_lines.py <src.csv from `ncu --page source --csv`> <nvdisasm -g output> <mangled function substring> [top]"""
import csv, re, sys, collections
src_csv, disasm, func = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# offset -> (file line)
line_of = {}
cur = None
infunc = False
for ln in open(disasm, errors="ignore"):
    if ln.startswith("\t.section") or ln.startswith(".section"):
        infunc = (".text." in ln and func in ln)
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
base = min(int(r[0], 16) for r in data)
agg = collections.defaultdict(lambda: collections.Counter())
def f(r, k):
    try:
        return float(r[idx[k]])
    except Exception:
        return 0.0
keys = ["# Samples", "Instructions Executed", "stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_no_inst",
        "stall_branch_resolving", "stall_mio", "stall_math", "stall_membar", "stall_not_selected", "stall_dispatch", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive"]
tot = collections.Counter()
for r in data:
    off = int(r[0], 16) - base
    loc = line_of.get(off, (None, ""))[0]
    for k in keys:
        agg[loc][k] += f(r, k)
        tot[k] += f(r, k)
print("total samples %d, warp instructions %d, shared wavefronts %d (excessive %d)" % (tot["# Samples"], tot["Instructions Executed"], tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Excessive"]))
srcs = {}
def text(loc):
    if not loc:
        return "?"
    fn, n = loc
    if fn not in srcs:
        import glob
        c = glob.glob("/root/repo/speech-intent-recognizer_b200/csrc/" + fn)
        srcs[fn] = open(c[0]).read().splitlines() if c else []
    return srcs[fn][n - 1].strip()[:90] if 0 < n <= len(srcs[fn]) else ""
print("%-22s %6s %6s | %5s %5s %5s %5s %5s %5s %5s | %6s %6s" % ("line", "smp%", "ins%", "lsb", "bar", "wait", "ssb", "noins", "brres", "mio", "shwf%", "exc%"))
for loc, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    name = "%s:%d" % loc if loc else "?"
    s = c["# Samples"] or 1
    print("%-22s %6.2f %6.2f | %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f %5.0f | %6.2f %6.2f  %s" % (
        name, 100 * c["# Samples"] / tot["# Samples"], 100 * c["Instructions Executed"] / tot["Instructions Executed"],
        100 * c["stall_long_sb"] / s, 100 * c["stall_barrier"] / s, 100 * c["stall_wait"] / s, 100 * c["stall_short_sb"] / s,
        100 * c["stall_no_inst"] / s, 100 * c["stall_branch_resolving"] / s, 100 * c["stall_mio"] / s,
        100 * c["L1 Wavefronts Shared"] / max(tot["L1 Wavefronts Shared"], 1), 100 * c["L1 Wavefronts Shared Excessive"] / max(tot["L1 Wavefronts Shared Excessive"], 1), text(loc)))
