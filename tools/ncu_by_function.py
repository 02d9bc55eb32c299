"""Per-device-function share of an ncu profile (samples, executed instructions, stall mix).

    python tools/ncu_by_function.py prof.ncu-rep libbsgp.so <kernel substring> 

Non-inlined device functions are subroutines inside the kernel's text section; nvdisasm labels them
`$kernel$function`.  Joins those label ranges with ncu's per-instruction SASS page.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, so, kern = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
func_of = {}
for cubin in [f for f in os.listdir(tmp) if f.endswith(".cubin")]:
    dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    inside, cur = False, "controller (kernel body)"
    for l in dis:
        if l.startswith(".text."):
            inside = kern in l
            cur = "controller (kernel body)"
            continue
        if not inside:
            continue
        m = re.match(r"^\s*\.type\s+(\$\S+),@function", l)
        if m:
            name = m.group(1).split("$")[-1]
            d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            mm = re.search(r"(?:bsgp::)?(\w+)\s*(?:<|\()", d)
            cur = mm.group(1) if mm else name
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
        if m:
            func_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(collections.Counter); tot = collections.Counter()
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    addr = int(r[col["Address"]], 16)
    if base is None: base = addr
    f = func_of.get(addr - base, "?")
    n = int(r[col["# Samples"]] or 0); ins = int(r[col["Instructions Executed"]] or 0)
    src = r[col["Source"]]
    agg[f]["samples"] += n; agg[f]["inst"] += ins; tot["samples"] += n; tot["inst"] += ins
    if "LDL" in src or "STL" in src: agg[f]["local"] += ins
    if re.search(r"\bD(ADD|MUL|FMA|SETP)\b|MUFU\.RCP64H", src): agg[f]["fp64"] += ins
    for s in stall_cols: agg[f][s] += int(r[col[s]] or 0)
print(f"total samples {tot['samples']}  warp instructions {tot['inst']}")
print(f"{'function':28s} {'time%':>6s} {'inst%':>6s} {'fp64%':>6s} {'local%':>6s}  top stalls")
for f, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    main = ", ".join(f"{s[6:]} {100 * c[s] // max(c['samples'], 1)}%" for s in sorted(stall_cols, key=lambda s: -c[s])[:3] if c[s])
    print(f"{f[:28]:28s} {100 * c['samples'] / tot['samples']:6.1f} {100 * c['inst'] / tot['inst']:6.1f} {100 * c['fp64'] / max(c['inst'], 1):6.1f} {100 * c['local'] / max(c['inst'], 1):6.1f}  {main}")
