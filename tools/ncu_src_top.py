"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv
import sys


def main():
    path, want = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    rows = list(csv.reader(open(path)))
    hdr, fpath, cur, items = None, "", ("", ""), []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if want and want not in fpath:
            continue
        if r[0]:
            cur = (fpath + ":" + r[0], r[1])
            continue
        if r[2] in ("", "..."):
            continue
        s = int(r[6]) if r[6].isdigit() else 0
        items.append((cur, r[3], s, r))
    tot = sum(i[2] for i in items)
    first_stall = hdr.index("# Samples") + 1
    names = [h for h in hdr]
    print("total samples", tot)
    for cur, sass, s, r in sorted(items, key=lambda x: -x[2])[:40]:
        st = sorted(((int(r[k]) if r[k].isdigit() else 0, names[k]) for k in range(32, len(hdr))), reverse=True)
        print(f"{s:6d} {100 * s / max(tot, 1):5.1f}% {cur[0]:>20s} {sass.strip()[:58]:58s} | {st[0][1][:24]}:{st[0][0]} | {cur[1].strip()[:44]}")


if __name__ == "__main__":
    main()
