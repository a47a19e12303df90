"""Instruction histogram of the Blackwell-specific SASS in libsscg_b200.so, per kernel (run on the build box:
cuobjdump needs no GPU).  UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, RED / ATOM = global reductions (integer only since round 2), HMMA = legacy
tensor path (must be absent).  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "semi-supervised-segmentation-cyclegan_b200", "libsscg_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "HGMMA",
        "RED.E", "REDG", "ATOMG", "ATOM.E", "ATOMS", "LDGSTS", "FFMA", "HFMA2", "DFMA", "DADD", "I2F.F64", "F2F.F32.F64"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void sscg::", "").replace("sscg::", "")
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k) or (k in ("RED.E", "ATOM.E") and op.startswith(k.split(".")[0] + ".")):
                    per[cur][k] += 1
            if op.startswith("RED") or op.startswith("ATOM"):
                per[cur]["atomics:" + op] += 1
    print("# SASS summary of libsscg_b200.so (cuobjdump -sass, sm_100a); counts are static instructions per kernel\n")
    tot = collections.Counter()
    for name, c in per.items():
        items = ["%s=%d" % (k, c[k]) for k in KEYS if c[k]]
        atom = ["%s x%d" % (k[8:], v) for k, v in c.items() if k.startswith("atomics:")]
        print("%-70s total=%-6d %s%s" % (name[:70], c["_total"], " ".join(items), ("   [" + ", ".join(atom) + "]") if atom else ""))
        tot.update({k: v for k, v in c.items() if not k.startswith("atomics:")})
        tot.update({k: v for k, v in c.items() if k.startswith("atomics:")})
    print("\n## whole library")
    print(" ".join("%s=%d" % (k, tot[k]) for k in KEYS if tot[k]))
    print("global / shared atomics by opcode: " + ", ".join("%s x%d" % (k[8:], v) for k, v in sorted(tot.items()) if k.startswith("atomics:")))
    fp_atomics = [k for k in tot if k.startswith("atomics:") and re.search(r"\.F32|\.F16|\.F64|\.BF16|FADD", k)]
    print("floating-point atomics: %s" % (", ".join(fp_atomics) if fp_atomics else "none"))


if __name__ == "__main__":
    main()
