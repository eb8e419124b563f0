"""Per-CUDA-source-line attribution of an ncu capture taken with -lineinfo / --import-source on.

    ncu -i gpurun_out/<report>.ncu-rep --page source --csv --print-source cuda,sass > /tmp/cs.csv
    python profiles/summarize_ncu_lines.py /tmp/cs.csv <units> [top] > profiles/<name>_lines.md
"""
import collections
import csv
import sys


def main(path, units, top=40):
    per, inst, src = collections.Counter(), collections.Counter(), {}
    stall = collections.defaultdict(collections.Counter)
    fname, hdr, ix = None, None, {}
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr, ix = r, {}
            for i, h in enumerate(hdr):
                ix.setdefault(h, i)
        elif hdr is not None and len(r) >= len(hdr) and r[0].isdigit() and r[2] == "-":
            key = (fname, int(r[0]))
            src[key] = r[1].strip()
            per[key] += int(r[ix["# Samples"]] or 0)
            inst[key] += int(r[ix["Instructions Executed"]] or 0)
            for h in ("stall_math", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_lg"):
                stall[key][h] += int(r[ix[h]] or 0)
    tot = sum(per.values())
    print("Samples attributed to source lines: %d.  Columns: share of samples; of which math (FP64 pipe busy), wait (fixed-latency,\n"
          "incl. the DMMA issue throttle), long_sb (mbarrier spin / global loads), short_sb (LDS, shuffles); warp instructions per unit.\n" % tot)
    print("| file:line | samples | math | wait | long_sb | short_sb | inst/unit | source |\n|---|---|---|---|---|---|---|---|")
    for k, s in per.most_common(top):
        st = stall[k]
        print("| %s:%d | %.1f%% | %.1f%% | %.1f%% | %.1f%% | %.1f%% | %.0f | `%s` |"
              % (k[0], k[1], 100.0 * s / tot, 100.0 * st["stall_math"] / tot, 100.0 * st["stall_wait"] / tot,
                 100.0 * st["stall_long_sb"] / tot, 100.0 * st["stall_short_sb"] / tot, inst[k] / units,
                 src[k][:90].replace("|", "\\|")))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40)
