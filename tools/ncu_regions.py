"""Region breakdown of an `ncu --page source --csv --print-source sass` export: consecutive SASS
instructions with about the same execution count are merged; prints instruction and stall-sample shares.
usage: python tools/ncu_regions.py export.csv [kernel-index ...]"""
import csv
import pickle
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
            continue
        if cur is None:
            continue
        if r and r[0] == "Address":
            cur["hdr"] = r
            continue
        if r:
            cur["rows"].append(r)
    return blocks


def main():
    blocks = load(sys.argv[1])
    which = [int(a) for a in sys.argv[2:]] or range(len(blocks))
    for bi in which:
        b = blocks[bi]
        h = b["hdr"]
        iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        out = [(i, int(r[iE]), int(r[iSm]), r[iS].strip()) for i, r in enumerate(b["rows"])]
        pickle.dump(out, open(f"/tmp/ncu_regions_{bi}.pkl", "wb"))
        tot = sum(e for _, e, _, _ in out)
        smp = sum(s for _, _, s, _ in out)
        print(b["name"][:60], "warp-instr", tot, "samples", smp, "sass", len(out))
        segs, c = [], None
        for i, e, s, src in out:
            if c and abs(e - c["e"]) <= 0.15 * max(e, c["e"], 1):
                c["n"] += 1; c["sum"] += e; c["smp"] += s; c["end"] = i
            else:
                c = {"start": i, "end": i, "e": e, "n": 1, "sum": e, "smp": s}
                segs.append(c)
        for sg in segs:
            if sg["sum"] > 0.006 * tot or sg["smp"] > 0.01 * smp:
                print(f"{sg['start']:5d}-{sg['end']:5d} n={sg['n']:4d} exec/instr={sg['e'] / 1e3:8.1f}K  sum={sg['sum'] / 1e6:6.2f}M "
                      f"({100 * sg['sum'] / tot:4.1f}%) samples={sg['smp']} ({100 * sg['smp'] / smp:4.1f}%)")


if __name__ == "__main__":
    main()
