"""SASS evidence for profiles/: which kernels of the shipped libpylc_b200.so use the Blackwell / Hopper+ data
movement (TMA: UTMALDG / UTMASTG, mbarrier: SYNCS.*, bulk-group waits), which use cp.async (LDGSTS), and
the loop of one TMA kernel around its load / store instructions.

    python tools/sass_excerpt.py [r2] > profiles/sass_r2.md        (no GPU needed: cuobjdump on the .so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pylc_b200", "lib", "libpylc_b200.so")
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
PAT = collections.OrderedDict([
    ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAPF (prefetch.tensormap)", r"\bUTMAPF|UTMACCTL"),
    ("UTMACMDFLUSH (bulk commit)", r"\bUTMACMDFLUSH"), ("SYNCS (mbarrier)", r"\bSYNCS\."), ("LDGSTS (cp.async)", r"\bLDGSTS"),
    ("DEPBAR (wait_group)", r"\bDEPBAR"), ("FFMA2 / FADD2 (f32x2, sm_100+)", r"\bFFMA2|\bFADD2|\bFMUL2"), ("REDUX", r"\bREDUX"),
    ("MUFU.EX2", r"MUFU\.EX2"), ("ATOMS", r"\bATOMS"),
])


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(line)
    print("# SASS of the shipped library (%s, sm_100a) -- %s" % (os.path.relpath(LIB, ROOT), R))
    print()
    print("`cuobjdump -sass` of `pylc_b200/lib/libpylc_b200.so`: %d device functions.  Counts are static instruction" % len(funcs))
    print("counts per kernel (PTX `cp.async.bulk.tensor` -> `UTMALDG` / `UTMASTG`, `mbarrier.*` -> `SYNCS.*`,")
    print("`cp.async` -> `LDGSTS`).  Only kernels with at least one of the listed instructions are shown.")
    print()
    keys = list(PAT)
    print("| kernel | " + " | ".join(keys) + " | instructions |")
    print("|---|" + "---|" * (len(keys) + 1))
    tot = collections.Counter()
    for f, lines in funcs.items():
        cnt = [sum(1 for l in lines if re.search(p, l)) for p in PAT.values()]
        for k, c in zip(keys, cnt):
            tot[k] += c
        if any(cnt[:6]):
            short = re.sub(r"\(.*", "", demangle(f)).replace("void ", "")
            print("| `%s` | " % short[:90] + " | ".join(str(c) if c else "" for c in cnt) + " | %d |" % len(lines))
    print()
    print("Totals over the library: " + ", ".join("%s %d" % (k, tot[k]) for k in keys) + ".")
    print()
    # excerpt: the TMA mask gather's main loop around its bulk tensor load / stores
    for f, lines in funcs.items():
        if "gather_mask_tma_kernel" in f and "Li5ELb1" in f:
            idx = [i for i, l in enumerate(lines) if re.search(r"UTMALDG|UTMASTG|SYNCS\.|UTMACMDFLUSH", l)]
            print("## Excerpt: `%s`" % re.sub(r"\(.*", "", demangle(f)))
            print()
            print("Every line that touches the copy engine or an mbarrier (static order; addresses in hex):")
            print()
            print("```")
            for i in idx:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", lines[i]).rstrip())
            print("```")
            break


if __name__ == "__main__":
    main()
