"""Per-source-line instruction and stall-sample totals from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > f.csv`.
    python profiles/src_lines.py f.csv [top_n]"""
import csv
import sys
from collections import OrderedDict


def main(fn, n=40):
    rows = list(csv.reader(open(fn)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    iline, isrc = 0, 1
    isamp = hdr.index("Warp Stall Sampling (All Samples)")
    iex = hdr.index("Instructions Executed")
    agg = OrderedDict()
    cur = None
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= iex:
            continue
        if r[iline].strip().isdigit():          # a CUDA source line; its SASS rows follow with an empty line number
            cur = (int(r[iline]), r[isrc].strip()[:100])
            agg.setdefault(cur, [0, 0])
            continue
        if cur is None:
            continue
        try:
            agg[cur][0] += int(r[iex] or 0)
            agg[cur][1] += int(r[isamp] or 0)
        except ValueError:
            pass
    tot_ex = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot_ex}, samples {tot_s}")
    for (ln, src), (ex, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
        print(f"{ln:5d} {100 * ex / tot_ex:5.1f}% inst {100 * s / tot_s:5.1f}% stall  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
