#!/usr/bin/env python
"""Per-instruction view of an .ncu-rep (source page, SASS): the instructions that carry the shared-memory wavefronts,
the global L1 tag requests and the stall samples. Usage: python tools/ncu_hot.py <file.ncu-rep> [top]"""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    def num(r, k):
        try: return float(r[col[k]].replace(",", ""))
        except Exception: return 0.0
    tot_inst = sum(num(r, "Instructions Executed") for r in body)
    tot_sh = sum(num(r, "L1 Wavefronts Shared") for r in body)
    tot_tag = sum(num(r, "L1 Tag Requests Global") for r in body)
    tot_samp = sum(num(r, "# Samples") for r in body)
    print("instructions %d  warp-inst executed %.3g  shared wavefronts %.3g  global tag requests %.3g  samples %d" % (len(body), tot_inst, tot_sh, tot_tag, tot_samp))
    for key in ("L1 Wavefronts Shared", "L1 Tag Requests Global", "# Samples"):
        print("== top by", key)
        for r in sorted(body, key=lambda r: -num(r, key))[:top]:
            print("  %6.2f%%  exec %.3g  %s" % (100 * num(r, key) / max(1.0, {"L1 Wavefronts Shared": tot_sh, "L1 Tag Requests Global": tot_tag, "# Samples": tot_samp}[key]),
                                               num(r, "Instructions Executed"), r[col["Source"]][:110]))
    ops = {}
    for r in body:
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"): op = r[col["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0.0) + num(r, "Instructions Executed")
    print("== executed warp instructions by opcode")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:30]:
        print("  %-10s %.3g (%.1f%%)" % (k, v, 100 * v / tot_inst))

if __name__ == "__main__":
    main()
