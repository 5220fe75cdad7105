"""SASS opcode counts per kernel of the built library -> profiles/r02_sass_summary.txt (run where cuobjdump is: here)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "conan-fgw_b200", "lib", "libconanmp.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "USETMAXREG", "SYNCS", "MUFU", "F2FP", "HFMA2", "FFMA"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = re.sub(r"^_ZN3cmp\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}\d+", "", m.group(1))
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        counts[name]["instr"] += 1
        for o in OPS:
            if op.startswith(o):
                counts[name][o] += 1
out = ["SASS opcode counts per kernel of conan-fgw_b200/lib/libconanmp.so (cuobjdump -sass, sm_100a; the .so is git-ignored,",
       "this file pins what the round-2 numbers were produced with; tools/sass_summary.py).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,",
       "UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, MUFU = special-function unit.", "",
       f"{'kernel':66s}{'instr':>7s}" + "".join(f"{o:>11s}" for o in OPS)]
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["instr"]):
    out.append(f"{k[:65]:66s}{c['instr']:7d}" + "".join(f"{c[o]:11d}" for o in OPS))
open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:16]))
