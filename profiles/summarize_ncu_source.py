"""Summarise the SASS source page of an ncu report: stall-reason shares, dynamic instruction mix, FP64-pipe time model.

    ncu -i gpurun_out/<report>.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv
    python profiles/summarize_ncu_source.py /tmp/sass.csv <units> > profiles/<name>.md

<units> = work units of the profiled launch (8-point tiles for the Burgers kernel, 16-point rounds x layers for NLS ...), used
only to print per-unit instruction counts.  Reads nothing but the CSV (runs on the CPU-only build container)."""
import collections
import csv
import re
import sys


def opcode(src):
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", src)
    return m.group(2) if m else "?"


def main(path, units):
    rows = list(csv.reader(open(path)))
    kernel = rows[0][1]
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) >= len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, per_stall, inst, samp = 0, collections.Counter(), collections.Counter(), collections.Counter()
    by_op_stall = collections.defaultdict(collections.Counter)
    for r in data:
        o = opcode(r[ix["Source"]])
        s = int(r[ix["# Samples"]] or 0)
        tot += s
        samp[o] += s
        inst[o] += int(r[ix["Instructions Executed"]] or 0)
        for c in stalls:
            v = int(r[ix[c]] or 0)
            per_stall[c] += v
            by_op_stall[c][o] += v
    n_inst = sum(inst.values())
    print("# ncu source-page summary: `%s`\n" % kernel)
    print("%d warp-state samples, %d warp-level instructions executed, %d work units.\n" % (tot, n_inst, units))
    print("| stall reason | share of samples | top opcodes |\n|---|---|---|")
    for c, v in per_stall.most_common(10):
        top = ", ".join("%s %.1f%%" % (o, 100.0 * n / tot) for o, n in by_op_stall[c].most_common(3))
        print("| %s | %.1f%% | %s |" % (c, 100.0 * v / tot, top))
    print("\n| opcode | executed | share | per unit | samples |\n|---|---|---|---|---|")
    for o, v in inst.most_common(18):
        print("| %s | %d | %.1f%% | %.0f | %.1f%% |" % (o, v, 100.0 * v / n_inst, v / units, 100.0 * samp[o] / tot))
    dmma = inst["DMMA"]
    fp64 = inst["DFMA"] + inst["DMUL"] + inst["DADD"]
    print("\nFP64-pipe time model per unit and SM sub-partition: DMMA %.0f x 16 = %.0f cycles, DFMA/DMUL/DADD %.0f x 2 = %.0f cycles, "
          "together %.0f cycles; all other instructions %.0f issue slots."
          % (dmma / units, 16.0 * dmma / units, fp64 / units, 2.0 * fp64 / units, (16.0 * dmma + 2.0 * fp64) / units,
             (n_inst - dmma - fp64) / units))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
