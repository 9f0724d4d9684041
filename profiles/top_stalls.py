"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv --kernel-name K`."""
import csv
import sys


def main(fn, n=24):
    rows = list(csv.reader(open(fn)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    isrc, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data, launch = [], 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            launch += 1
        if launch != 1 or len(r) <= max(isamp, iex) or not r[isamp].isdigit():
            continue
        data.append((int(r[isamp]), r[isrc].strip(), r[iex], len(data)))
    tot = sum(d[0] for d in data)
    print("total samples", tot, "instructions", len(data))
    for t in sorted(data, reverse=True)[:n]:
        print(f"{t[0]:7d} {100 * t[0] / tot:5.1f}%  #{t[3]:4d} ex={t[2]:>9}  {t[1][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
