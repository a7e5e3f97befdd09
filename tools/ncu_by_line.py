"""Join an ncu SASS-level source page (ncu -i rep --page source --csv) with nvdisasm --print-line-info of the same
cubin: samples / executed instructions per source line and per source function range.
usage: python tools/ncu_by_line.py src.csv lines.sass kernel_symbol [source.cu]"""
import csv
import re
import sys
from collections import defaultdict

csv.field_size_limit(10**9)


def main():
    src_csv, sass, sym = sys.argv[1:4]
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    base = None
    prof = {}
    for r in rows[2:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        prof[a - base] = (int(r[isamp] or 0), int(r[iex] or 0), [int(r[i] or 0) for i in stall_cols], r[1].strip())
    # line info
    cur = None
    infn = False
    amap = {}
    for ln in open(sass):
        if ln.startswith(".text." + sym):
            infn = True
            continue
        if infn and ln.startswith(".text.") and sym not in ln:
            break
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/", ln)
        if m:
            amap[int(m.group(1), 16)] = cur
    by_line = defaultdict(lambda: [0, 0, 0])
    stall_by_line = defaultdict(lambda: [0] * len(stall_cols))
    for off, (s, ex, st, txt) in prof.items():
        k = amap.get(off, ("?", 0))
        by_line[k][0] += s
        by_line[k][1] += ex
        by_line[k][2] += 1
        for i, v in enumerate(st):
            stall_by_line[k][i] += v
    tot_s = sum(v[0] for v in by_line.values())
    tot_e = sum(v[1] for v in by_line.values())
    print("total samples %d, instructions executed %d, static instructions %d" % (tot_s, tot_e, len(prof)))
    names = [hdr[i] for i in stall_cols]
    print("top lines by samples:")
    for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:60]:
        st = stall_by_line[k]
        top = sorted(zip(st, names), reverse=True)[:3]
        print("  %-22s:%-5d samples %6.2f%%  exec %6.2f%%  static %5d   %s" % (k[0], k[1], 100.0 * v[0] / tot_s, 100.0 * v[1] / tot_e, v[2],
                                                                  " ".join("%s=%.0f%%" % (n[6:], 100.0 * s / max(1, sum(st))) for s, n in top)))
    # per file ranges of 1 function: caller passes ranges "file:lo-hi:name"
    for spec in sys.argv[4:]:
        f, rng, name = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        s = sum(v[0] for k, v in by_line.items() if k[0] == f and lo <= k[1] <= hi)
        e = sum(v[1] for k, v in by_line.items() if k[0] == f and lo <= k[1] <= hi)
        n = sum(v[2] for k, v in by_line.items() if k[0] == f and lo <= k[1] <= hi)
        st = [sum(stall_by_line[k][i] for k in by_line if k[0] == f and lo <= k[1] <= hi) for i in range(len(names))]
        top = sorted(zip(st, names), reverse=True)[:4]
        print("%-28s samples %6.2f%% exec %6.2f%% static %6d  %s" % (name, 100.0 * s / tot_s, 100.0 * e / tot_e, n,
                                                                  " ".join("%s=%.0f%%" % (nn[6:], 100.0 * x / max(1, sum(st))) for x, nn in top)))


if __name__ == "__main__":
    main()
