"""Per-source-line summary of an ncu source page: joins the SASS-level metrics of one kernel with the
line table of the cubin (nvdisasm -g) by instruction order.
    python scripts/ncu_by_line.py report.ncu-rep object.o kernel_substring [source.cu] [top]"""
import csv, re, subprocess, sys, os, tempfile, collections
rep, obj, kern = sys.argv[1:4]
src = sys.argv[4] if len(sys.argv) > 4 else None
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel with their current source line
lines, cur, inside = [], None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# split per kernel
blocks, curb = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curb = {"name": r[1], "rows": []}; blocks.append(curb)
    elif curb is not None:
        curb["rows"].append(r)
b = [x for x in blocks if kern.replace("ILi", "<").split("<")[0].split("8k_")[-1] in x["name"] or True][0]
hdr = b["rows"][0]; data = b["rows"][1:]
iS, iE, iT, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
print("kernel:", b["name"], "| sass rows", len(data), "| line-table instructions", len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
n = min(len(data), len(lines))
for k in range(n):
    r = data[k]
    v = [int(r[iS] or 0), int(r[iE] or 0), int(r[iT] or 0)]
    key = lines[k]
    for j in range(3):
        agg[key][j] += v[j]; tot[j] += v[j]
print("total: stall samples %d, warp instr %d, thread instr %d (avg %.1f lanes)" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
text = {}
if src:
    for i, l in enumerate(open(src), 1):
        text[i] = l.rstrip()
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = key if key else ("?", 0)
    print("%5.1f%% samples %5.1f%% instr  lanes %4.1f  %s:%d  %s" % (100.0 * v[0] / max(tot[0], 1), 100.0 * v[1] / max(tot[1], 1),
          v[2] / max(v[1], 1), f, ln, text.get(ln, "")[:110] if f == os.path.basename(src or "") else ""))
