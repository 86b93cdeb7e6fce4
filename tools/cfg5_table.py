"""Merge the per-G sweep files of BASELINE configs[4] and print the markdown table of BASELINE.md section 5.
usage: python tools/cfg5_table.py [--write]   (--write: rewrite profiles/r2_cfg5.jsonl and the table in BASELINE.md)"""
import json, os, re, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for G in (1, 2, 4, 8):
    path = os.path.join(root, "profiles", f"r2_cfg5_g{G}.jsonl")
    rows += [json.loads(l) for l in open(path) if l.strip().startswith("{")]
gpu, cpu = {}, {}
for r in rows:
    if "skipped" in r:
        continue
    mm = re.search(r"(linear\+icpt|linear) n=(\d+) f=(\d+) m=([0-9e+.]+) G=(\d+)", r["config"])
    fam, n, m, G = mm.group(1), int(mm.group(2)), float(mm.group(4)), int(mm.group(5))
    if "cpu" in r:
        cpu[(fam, n, m, r["kernel"], r["cpu"].split("[")[0])] = r
    else:
        gpu[(fam, n, m, G, r["kernel"])] = r
def fm(m):
    e = len("%d" % m) - 1
    c = m / 10 ** e
    return ("%d·10^%d" % (c, e)) if c != 1 else "10^%d" % e
out = ["| family, n | kernel | m | G=1 | G=2 | G=4 | G=8 | CPU seq (1 core) | CPU par[16] |", "|---|---|---|---|---|---|---|---|---|"]
for fam, n in sorted({(k[0], k[1]) for k in gpu}, key=lambda t: (t[1], t[0])):
    for kern in ("K1", "K2"):
        for m in sorted({k[2] for k in gpu}):
            cells = []
            for G in (1, 2, 4, 8):
                r = gpu.get((fam, n, m, G, kern))
                cells.append("%.3f / %.3f" % (r["wall_ms_per_call"], r["kernel_ms"]) if r else "—")
            if all(c == "—" for c in cells):
                continue
            cs, cp = cpu.get((fam, n, m, kern, "seq")), cpu.get((fam, n, m, kern, "par"))
            out.append(f"| {fam} n={n} | {kern} | {fm(m)} | " + " | ".join(cells) +
                       " | %s | %s |" % ("%.3g" % (m / cs["rows_per_s"] * 1e3) if cs else "", "%.3g" % (m / cp["rows_per_s"] * 1e3) if cp else ""))
table = "\n".join(out) + "\n"
if "--write" in sys.argv:
    with open(os.path.join(root, "profiles", "r2_cfg5.jsonl"), "w") as fh:
        for r in rows:
            fh.write(json.dumps(r) + "\n")
    p = os.path.join(root, "BASELINE.md")
    s = open(p).read()
    a = s.index("| family, n | kernel | m |")
    open(p, "w").write(s[:a] + table)
else:
    print(table, end="")
