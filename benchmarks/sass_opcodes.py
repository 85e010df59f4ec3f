"""profiles/r02_sass_opcodes.md: per-kernel counts of the Blackwell-specific SASS opcodes in the built library
(`cuobjdump -sass pvnet_b200/_lib/libpvnet_b200.so`).  Runs in the authoring container (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pvnet_b200", "_lib", "libpvnet_b200.so")
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "FFMA2", "FMNMX3", "LEA.HI", "REDG", "LDS.128",
       "DSETP", "DFMA"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.md")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    rows, tot = [], collections.Counter()
    for f in funcs[1:]:
        mangled = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        m = re.search(r"(k_\w+)(<[^(]*?>)?\(", dem)
        short = (m.group(1) + (m.group(2) or "")) if m else dem[:60]
        short = short.replace("(bool)", "").replace("(int)", "")
        c = {o: len(re.findall(r"\s" + re.escape(o) + r"[\s.]", f)) for o in OPS}
        n = len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+[A-Z@]", f))
        rows.append((short, n, c))
        for o in OPS:
            tot[o] += c[o]
    lines = ["# SASS opcode counts of pvnet_b200/_lib/libpvnet_b200.so (sm_100a), per kernel", "",
             "`python benchmarks/sass_opcodes.py` = `cuobjdump -sass` of the in-tree library, occurrences per entry point.",
             "UTCHMMA = tcgen05.mma; UTMALDG / UTMASTG = TMA load / store (cp.async.bulk.tensor); LDTM / STTM = tcgen05.ld / .st",
             "(TMEM); UTCBAR = tcgen05.commit; SYNCS = mbarrier; ELECT = elect.sync; FFMA2 = packed fp32 FMA (fma.rn.f32x2) and",
             "FFMA.SAT = count-by-saturation, both in the vote kernel; REDG = the one global reduction per hypothesis; DFMA/DSETP =",
             "fp64 (refit sums, PnP, the reference's `< 1e-6` tests in double).", "",
             "| kernel | instructions | " + " | ".join(OPS) + " |", "|---|---|" + "---|" * len(OPS)]
    for short, n, c in sorted(rows, key=lambda r: -r[1]):
        lines.append(f"| `{short}` | {n} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    lines.append("| **total** | | " + " | ".join(str(tot[o]) for o in OPS) + " |")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print(f"{len(rows)} kernels ->", out_path)


if __name__ == "__main__":
    main()
