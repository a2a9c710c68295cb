"""Aggregate an `ncu --page source --csv` (SASS view) export into regions delimited by barrier instructions, so the
per-phase share of stall samples / executed instructions of a multi-phase kernel can be read without the GUI."""
import csv
import sys


def main(path, extra_markers=()):
    rows = list(csv.reader(open(path)))
    # keep only the first kernel block of the export
    ends = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if len(ends) > 1:
        rows = rows[: ends[1]]
    hdr = rows[1]
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    regions, cur = [], {"start": 0, "samples": 0, "inst": 0, "n": 0, "marker": "entry", "top": []}
    for k, r in enumerate(rows[2:]):
        if len(r) <= max(si, ii):
            continue
        s, i, txt = int(r[si] or 0), int(r[ii] or 0), r[ti].strip()
        cur["samples"] += s; cur["inst"] += i; cur["n"] += 1
        cur["top"].append((s, txt))
        if any(m in txt for m in ("BAR.", "BAR ", "EXIT") + tuple(extra_markers)):
            regions.append(cur)
            cur = {"start": k + 1, "samples": 0, "inst": 0, "n": 0, "marker": txt, "top": []}
    regions.append(cur)
    tot_s = sum(r["samples"] for r in regions) or 1
    tot_i = sum(r["inst"] for r in regions) or 1
    print(f"total samples {tot_s}, total warp-instructions {tot_i}")
    for r in regions:
        if r["n"] == 0:
            continue
        top = sorted(r["top"], reverse=True)[:3]
        print(f"region@{r['start']:5d} n={r['n']:4d} samples={100*r['samples']/tot_s:5.1f}% inst={100*r['inst']/tot_i:5.1f}%  after [{r['marker'][:40]}]")
        for s, t in top:
            if s:
                print(f"        {100*s/tot_s:4.1f}%  {t[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
