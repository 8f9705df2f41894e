"""Condenses tools/perf_map.py outputs into profiles/r2/perf_map_table.txt:
    python tools/perf_map_table.py base.txt final.txt spatial_first_v1.txt spatial_first_final.txt average_v1.txt average_final.txt
`start > end` per cell; r = TMA row kernel, f = flex kernel, g = gather kernel, p = pooling kernel."""
import statistics
import sys


def load(path):
    d = {}
    for line in open(path):
        t = line.split()
        if len(t) >= 7 and "x" in t[0] and t[1].startswith("f="):
            d[(t[0], int(t[1][2:]), int(t[2][4:]), t[3], int(t[4][5:]))] = (t[5], float(t[6]))
    return d


def main():
    base, final = load(sys.argv[1]), load(sys.argv[2])
    sizes = []
    for k in final:
        if k[0] not in sizes:
            sizes.append(k[0])
    print("size      | f=1: YCC888 BUNDLE128 RGB888(6/5/5) | f=2 | f=4 | f=8")
    for s in sizes:
        row = f"{s:10s}"
        for f in (1, 2, 4, 8):
            for fmt in (0, 3, 1):
                k = (s, f, fmt, "CSQ", 0)
                b, e = base.get(k), final.get(k)
                row += f" {e[0][0]}{b[1]:.2f}>{e[1]:.2f}" if b else f" {e[0][0]}{e[1]:.2f}"
            row += " |"
        print(row)
    for title, v1, fin in (("spatial -> colour -> chroma (SQC; case B)", 3, 4), ("AVERAGE extension, CSQ and SQC", 5, 6)):
        if len(sys.argv) <= fin:
            break
        a, b = load(sys.argv[v1]), load(sys.argv[fin])
        print(f"\n{title}: {len(b)} configurations")
        for fam in sorted({v[0] for v in b.values()}):
            now = [v[1] for v in b.values() if v[0] == fam]
            was = [a[k][1] for k, v in b.items() if v[0] == fam and k in a]
            print(f"  {fam:8s} n={len(now):3d}  min {min(now):.2f}  median {statistics.median(now):.2f}  max {max(now):.2f}"
                  + (f"   (before: min {min(was):.2f} median {statistics.median(was):.2f})" if was else ""))


if __name__ == "__main__":
    main()
