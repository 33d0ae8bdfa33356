"""Where the spill instructions of a kernel sit relative to its loops: lists LDL / STL / ATOMG / LDTM addresses and every backward
branch of one function of the shipped library (cuobjdump -sass).  usage: sass_loop_check.py <lib.so> <mangled-name substring>"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, cur = {}, None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); fn[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,}\*/", ln):
        fn[cur].append(ln)
for name, lines in fn.items():
    if pat not in name:
        continue
    ins = []
    for ln in lines:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name, len(ins), "instructions")
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?(?:0x)?([0-9a-f]+)\)?", t)
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
    for lo, hi in sorted(loops):
        body = [t for a, t in ins if lo <= a <= hi]
        ops = {}
        for t in body:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:10]
        print("  loop %05x..%05x  %4d instr  LDL %d STL %d  | %s" % (lo, hi, len(body), ops.get("LDL", 0), ops.get("STL", 0), " ".join("%s:%d" % kv for kv in top)))
