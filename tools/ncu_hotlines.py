"""Per-source-line stall summary of a kernel from an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_hotlines.py gpurun_out/prof.ncu-rep beta-sgp_b200/csrc/libbsgp.so bsgp_solve_kernelIdLi512ELi1E [top]

Joins ncu's SASS page (stall samples per instruction) with nvdisasm -g line info of the cubin.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of = {}
inside, cur = False, None
for l in dis:
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[col["Address"]], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    n = int(r[col["# Samples"]] or 0)
    agg[key]["samples"] += n
    agg[key]["inst"] += int(r[col["Instructions Executed"]] or 0)
    tot["samples"] += n
    for s in stall_cols:
        v = int(r[col[s]] or 0)
        agg[key][s] += v
        tot[s] += v
print("total samples", tot["samples"], {s: tot[s] for s in stall_cols if tot[s] > tot["samples"] * 0.01})
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    main = ", ".join(f"{s[6:]} {c[s]}" for s in sorted(stall_cols, key=lambda s: -c[s])[:3] if c[s])
    print(f"{key[0]}:{key[1]:<5d} samples {c['samples']:8d} ({100.0 * c['samples'] / max(tot['samples'], 1):5.1f}%) inst {c['inst']:12d}  {main}")
