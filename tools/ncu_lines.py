"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
executed warp instructions and stall samples, top N.  Usage: ncu_lines.py src.csv [N]"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, cur_fn, hdr = None, None, None
    agg = defaultdict(lambda: [0, 0, ""])  # (file, line) -> [inst, samples, text]
    stall_cols = {}
    stalls = defaultdict(lambda: defaultdict(int))
    fn_tot = defaultdict(lambda: [0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            continue
        if hdr is None or r[0] == "":
            continue  # SASS rows
        try:
            line = int(r[0])
        except ValueError:
            continue
        if len(r) != len(hdr):
            continue  # a source line whose text broke the CSV quoting (inline asm with quotes): a handful of intrinsics
        try:
            inst = int(r[hdr.index("Instructions Executed")] or 0)
            smp = int(r[hdr.index("# Samples")] or 0)
        except ValueError:
            continue
        key = (cur_file, line)
        agg[key][0] += inst
        agg[key][1] += smp
        agg[key][2] = r[1].strip()[:110]
        fn_tot[cur_fn.split("(")[0][-40:]][0] += inst
        fn_tot[cur_fn.split("(")[0][-40:]][1] += smp
        for i, h in stall_cols.items():
            v = int(r[i] or 0)
            if v:
                stalls[key][h] += v
    tot_i = sum(v[0] for v in agg.values())
    tot_s = sum(v[1] for v in agg.values())
    print("total warp instructions %.3e, samples %d" % (tot_i, tot_s))
    print("-- per function")
    for k, v in sorted(fn_tot.items(), key=lambda kv: -kv[1][1]):
        print("  %-42s inst %5.1f%%  samples %5.1f%%" % (k, 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s))
    print("-- top lines by stall samples")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = sorted(stalls[key].items(), key=lambda kv: -kv[1])[:3]
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s   [%s]" % (100.0 * v[1] / tot_s, 100.0 * v[0] / tot_i, key[0], key[1], v[2],
                                                            ", ".join("%s %d" % (a.replace("stall_", ""), b) for a, b in st)))
    allst = defaultdict(int)
    for d in stalls.values():
        for h, v in d.items():
            allst[h] += v
    print("-- stall reasons overall")
    for h, v in sorted(allst.items(), key=lambda kv: -kv[1])[:10]:
        print("  %-28s %5.1f%%" % (h, 100.0 * v / max(1, sum(allst.values()))))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)


def ranges(path, spec):
    """spec: list of (name, file, lo, hi) -> instruction / sample share per source range."""
    rows = list(csv.reader(open(path)))
    cur_file, hdr = None, None
    tot_i = tot_s = 0
    acc = {name: [0, 0] for name, _, _, _ in spec}
    other = defaultdict(lambda: [0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("", "Function Name"):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        if len(r) != len(hdr):
            continue
        try:
            inst = int(r[hdr.index("Instructions Executed")] or 0)
            smp = int(r[hdr.index("# Samples")] or 0)
        except ValueError:
            continue
        tot_i += inst
        tot_s += smp
        hit = False
        for name, f, lo, hi in spec:
            if cur_file == f and lo <= line <= hi:
                acc[name][0] += inst
                acc[name][1] += smp
                hit = True
                break
        if not hit:
            other[cur_file][0] += inst
            other[cur_file][1] += smp
    for name, _, _, _ in spec:
        print("  %-28s inst %5.1f%%  samples %5.1f%%" % (name, 100.0 * acc[name][0] / tot_i, 100.0 * acc[name][1] / tot_s))
    for f, v in sorted(other.items(), key=lambda kv: -kv[1][0]):
        print("  (other) %-20s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s))
