"""Turn the raw ncu outputs in gpurun_out/ into the small, tracked summaries under profiles/:
  profiles/launches_<round>.csv        per-kernel totals of the launch list (time share of a bench step)
  profiles/ncu_<round>_summary.md      key metrics of every `--set full` capture
Run here (no GPU needed): python tools/summarise_profiles.py r1
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


def launch_list():
    path = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % R)
    if not os.path.isfile(path):
        return
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    unit = hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[unit], 1.0)
        tot[short(r[kn])] += v
        cnt[short(r[kn])] += 1
    total = sum(tot.values())
    ours = ("pylc::",)
    with open(os.path.join(OUT, "launches_%s.csv" % R), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 1 --warmup 1 --images 2`\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolute times. total_us=%.1f launches=%d\n" % (total, sum(cnt.values())))
        f.write("kernel,launches,total_us,share,ours\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write('"%s",%d,%.1f,%.4f,%d\n' % (k, cnt[k], tot[k], tot[k] / total, int(any(o in k for o in ours))))
    mine = sum(v for k, v in tot.items() if any(o in k for o in ours))
    print("launch list: %d launches, %.1f ms total, custom kernels %.2f%%" % (sum(cnt.values()), total / 1e3, 100 * mine / total))


KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def captures():
    reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_%s_*.ncu-rep" % R)))
    if not reps:
        return
    traffic = {}
    md = ["# ncu `--set full --clock-control none` captures, round %s" % R, "",
          "One launch per kernel at the BASELINE sizes of `tools/kbench.py` (6000x4000 image, 273 logit tiles,",
          "B=64 loss batch), after 3 warm-up launches.  ncu flushes caches and serialises, so durations are",
          "cold-cache; the CUDA-event numbers in `kbench_%s.jsonl` are the ones quoted as roofline fractions." % R, ""]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            md.append("## `%s`  (%s)" % (name, os.path.basename(rep)))
            md.append("")
            md.append("| metric | value |")
            md.append("|---|---|")
            try:     # dram bytes per launch, keyed by capture name (bench.py reads traffic_<round>.json)
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                tag = os.path.basename(rep)[len("prof_%s_" % R):-len(".ncu-rep")]
                traffic.setdefault(tag, []).append({"kernel": name, "dram_bytes": float(r[rd]) * mult[units[rd]] + float(r[wr]) * mult[units[wr]]})
            except (ValueError, KeyError):
                pass
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    md.append("| %s | %s %s |" % (label, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
                    try:
                        stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            tot = sum(s[0] for s in stalls) or 1.0
            md.append("| top stall reasons (pc samples) | %s |" % ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in stalls[:4]))
            md.append("")
    with open(os.path.join(OUT, "ncu_%s_summary.md" % R), "w") as f:
        f.write("\n".join(md) + "\n")
    import json
    with open(os.path.join(OUT, "traffic_%s.json" % R), "w") as f:
        json.dump(traffic, f, indent=1)
    print("captures:", len(reps))


launch_list()
captures()
