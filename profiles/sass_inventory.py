#!/usr/bin/env python
"""SASS inventory of the built library: per kernel, how often the mnemonics occur that show which hardware path it uses.

  python profiles/sass_inventory.py dvmslam_b200/lib/libdvmslam_b200.so profiles/<name>.md

Runs where nvcc's tools are (no GPU needed)."""
import collections
import re
import subprocess
import sys

SPECIAL = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "STAS", "UCGABAR", "DMMA", "ELECT", "IDP"]
OTHER = ["ATOMG", "ATOMS", "RED", "BAR.SYNC", "BAR.RED", "BAR.ARV", "CREDUX", "REDUX", "MATCH", "DADD", "DFMA", "DMUL", "LDG", "LDS", "LDL", "STG",
         "STS", "STL", "MUFU", "POPC", "PRMT", "SHFL"]


def main(so, out):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_n"] += 1
            for k in SPECIAL + OTHER:
                if op == k or op.startswith(k + ".") or op.startswith(k + "_") or (k in ("BAR.SYNC", "BAR.RED", "BAR.ARV") and op.startswith(k)):
                    cur[k] += 1
                    break
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    with open(out, "w") as f:
        f.write("# SASS inventory of the built library\n\n`cuobjdump -sass " + so + "` (sm_100a), made by `profiles/sass_inventory.py`: per kernel, how often the "
                "mnemonics occur that show which hardware path it uses.\n`UTCIMMA` = tcgen05.mma kind::i8, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit, "
                "`UTMALDG` = TMA tensor load, `SYNCS` = mbarrier, `STAS` = st.async to a peer CTA's shared memory, `UCGABAR` = cluster barrier, "
                "`DMMA` = FP64 tensor-core MMA, `IDP` = integer dot product (dp2a / dp4a), `PRMT` = byte permute.\n\n"
                "| kernel | instructions | tensor / TMA / TMEM / cluster / dot | other |\n|---|---|---|---|\n")
        for (name, c), dn in sorted(zip(kernels.items(), demangled), key=lambda x: x[1]):
            short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "")).replace("dvm::", "").replace("void ", "")
            sp = ", ".join(f"{k} {c[k]}" for k in SPECIAL if c[k]) or "-"
            ot = ", ".join(f"{k} {c[k]}" for k in OTHER if c[k])
            f.write(f"| `{short}` | {c['_n']} | {sp} | {ot} |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
