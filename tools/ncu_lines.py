"""Per-source-line sample / instruction totals of one kernel from an ncu report with --import-source on.
The SASS page (ncu --page source --csv) is joined with the line table of the locally built cubin (nvdisasm -g), which is the
same binary the GPU box ran.  usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]"""
import csv, io, re, subprocess, sys
rep, cubin, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.split("\n")
seq, cur, on = [], None, False
for l in dis:
    if l.startswith(".text."):
        on = pat in l
        continue
    if not on:
        continue
    m = re.search(r'## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h], rows[h + 1:]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
assert len(data) == len(seq), (len(data), len(seq))
by = {}
tot = toti = 0
for (ln, _), r in zip(seq, data):
    s_, n_ = int(r[si] or 0), int(r[ii] or 0)
    tot += s_; toti += n_
    e = by.setdefault(ln, [0, 0]); e[0] += s_; e[1] += n_
print("samples", tot, "instructions", toti)
for ln, (s_, n_) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"line {ln}: samples {s_} ({100.0 * s_ / tot:.1f}%)  instr {n_} ({100.0 * n_ / toti:.1f}%)")
