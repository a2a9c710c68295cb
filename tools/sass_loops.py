"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with their static size and instruction mix:
a quick estimate of the per-point instruction count of solve_kernel's sweep before spending GPU time.
usage: cuobjdump -sass file.o | python tools/sass_loops.py <kernel-name-substring>"""
import collections
import re
import sys

pat = sys.argv[1]
lines = sys.stdin.read().splitlines()
cur, body = None, []
for ln in lines:
    if "Function :" in ln:
        cur = ln.split("Function :")[1].strip()
        continue
    if cur and pat in cur:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            body.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(body)}
for i, (a, ins) in enumerate(body):
    m = re.search(r"BRA\s+(?:\w+,\s*)?(0x[0-9a-f]+)", ins)
    if m:
        t = int(m.group(1), 16)
        if t < a and t in addr:
            seg = body[addr[t]: i + 1]
            ops = collections.Counter()
            for _, s in seg:
                toks = s.split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                ops[op.split(".")[0]] += 1
            if len(seg) >= int(sys.argv[2]) if len(sys.argv) > 2 else ops.get("DFMA", 0) >= 6:
                print(f"loop {t:#x}..{a:#x}: {len(seg)} instr;", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
