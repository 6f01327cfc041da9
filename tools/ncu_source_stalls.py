"""Aggregate the warp-stall samples of an `ncu --page source --csv` export per kernel (evidence for profiles/)."""
import collections
import csv
import gzip
import sys


def main(path, topn=14):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kern.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = r
            else:
                cur["rows"].append(r)
    seen = set()
    for k in kern:
        nm = k["name"][:70]
        if nm in seen:
            continue
        seen.add(nm)
        h = k["hdr"]
        idx = {n: i for i, n in enumerate(h)}
        stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = collections.Counter()
        for r in k["rows"]:
            for c in stall_cols:
                try:
                    tot[c] += int(r[idx[c]])
                except (ValueError, IndexError):
                    pass
        s = sum(tot.values()) or 1
        print(f"## {nm}  ({s} samples)")
        print("   " + ", ".join(f"{c[6:]}={100 * v / s:.1f}%" for c, v in tot.most_common(9)))
        top = sorted(k["rows"], key=lambda r: -int(r[idx["# Samples"]] or 0))[:topn]
        for r in top:
            print(f"      {r[idx['# Samples']]:>7}  {r[1].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
