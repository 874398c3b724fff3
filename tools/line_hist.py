"""Map the per-SASS-instruction counts of an `ncu --page source --csv` export onto CUDA source lines, using the
line info of `nvdisasm -g -c` for the same kernel.
usage: line_hist.py <ncu_source.csv> <nvdisasm.txt> <mangled kernel substring> [top_n]"""
import csv, collections, re, sys

src_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# offsets -> (file, line)
line_of = {}
cur = None
in_k = False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        in_k = kern in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
base = None
agg = collections.Counter()
smp = collections.Counter()
tot = stot = 0
for r in rows[h + 1 :]:
    if len(r) < len(hdr):
        continue
    try:
        addr = int(r[0], 16)
        n = int(r[idx["Instructions Executed"]] or 0)
        s = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    agg[key] += n
    smp[key] += s
    tot += n
    stot += s
print(f"total warp instr {tot}, samples {stot}, mapped lines {len(agg)}")
for key, n in agg.most_common(top):
    print(f"{n:12d} {100*n/tot:5.1f}%  smp {100*smp[key]/max(stot,1):5.1f}%  {key[0]}:{key[1]}")
