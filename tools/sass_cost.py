"""Issue-cost estimate of a straight-line SASS region (tools/k2_ab.sh variants): FP64 instructions cost 2 cycles (3 when they
read three distinct registers none of which is served by the operand reuse cache), everything else 1 (tools/dfma_mix.cu).
usage: python tools/sass_cost.py <exe or .o> <kernel-name-regex> <start-hex> <end-hex>"""
import re, subprocess, sys
exe, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", exe], capture_output=True, text=True).stdout
on = False; rows = []
for line in txt.splitlines():
    if "Function :" in line: on = re.search(pat, line) is not None
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if on and m:
        a = int(m.group(1), 16)
        if lo <= a < hi: rows.append((a, m.group(2).strip()))
fp64 = other = three = 0; prev_reuse = {}
ops = {}
for a, ins in rows:
    ins = re.sub(r"^@!?U?P\d\s+", "", ins)
    op = ins.split()[0]
    ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
    if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        fp64 += 1
        srcs = [s.strip() for s in ins[len(op):].split(",")][1:]
        reads = set(); new_reuse = {}
        for slot, s in enumerate(srcs):
            m = re.match(r"[-|]*R(\d+)(\.reuse)?", s)
            if not m: continue
            r = m.group(1)
            if m.group(2): new_reuse[slot] = r
            if prev_reuse.get(slot) == r: continue
            reads.add(r)
        if len(reads) >= 3: three += 1
        prev_reuse = new_reuse
    else:
        other += 1
print({"fp64": fp64, "three_read": three, "other": other, "cycles": 2 * fp64 + three + other, "ops": dict(sorted(ops.items(), key=lambda kv: -kv[1]))})
