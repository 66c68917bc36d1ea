"""Static evidence for profiles/: which Blackwell instructions each kernel of libp2r_b200.so contains and what it costs
in registers / shared memory (run here, on the CPU box: `python tools/sass_evidence.py profiles/r02_sass_evidence.txt`).

Per kernel (cuobjdump -sass / -res-usage of the shipped sm_100a library): counts of UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (tensor TMA load / store / reduce), UBLKCP (1-D bulk TMA), LDGSTS
(cp.async), SYNCS (mbarrier), legacy HMMA (must be 0), plus REG / SHARED / LOCAL(spill) from the resource table.
The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md lists as proof of a Blackwell-native kernel."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pose2room_b200", "lib", "libp2r_b200.so")
CLASSES = [("UTCxMMA", r"\bUTC[A-Z]*MMA"), ("2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
           ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"), ("UBLKCP", r"\bUBLKCP"),
           ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"(?<![A-Z])HMMA")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?!anonymous).*$", "", name)
    return name[:96]


def main(dst):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
        usage[m.group(1)] = m.group(2)
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = dict((k, 0) for k, _ in CLASSES)
            kernels[cur]["instr"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[-1] if line.lstrip().startswith("/*") else line
        if re.search(r"\b[A-Z][A-Z0-9_.]+\b", body) and ";" in body:
            kernels[cur]["instr"] += 1
            for k, pat in CLASSES:
                if re.search(pat, body):
                    kernels[cur][k] += 1
    names = demangle(list(kernels))
    rows = []
    for k, c in kernels.items():
        u = usage.get(k, "")
        reg = re.search(r"REG:(\d+)", u)
        sh = re.search(r"SHARED:(\d+)", u)
        loc = re.search(r"LOCAL:(\d+)", u)
        rows.append((short(names.get(k, k)), c, reg.group(1) if reg else "?", sh.group(1) if sh else "?",
                     loc.group(1) if loc else "?"))
    rows.sort(key=lambda r: (-r[1]["UTCxMMA"], -r[1]["UTMALDG"] - r[1]["UBLKCP"], r[0]))
    cols = [k for k, _ in CLASSES]
    with open(dst, "w") as f:
        f.write("# static SASS / resource evidence of pose2room_b200/lib/libp2r_b200.so (sm_100a), %d kernels\n" % len(rows))
        f.write("# command: python tools/sass_evidence.py %s   (cuobjdump -sass / -res-usage, CUDA 12.9)\n" % os.path.relpath(dst, ROOT))
        f.write("# UTCxMMA = tcgen05.mma, 2CTA = its cta_group::2 form, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor\n")
        f.write("# load/store/reduce-add, UBLKCP = 1-D bulk TMA, LDGSTS = cp.async, SYNCS = mbarrier ops, HMMA = legacy mma.sync (none)\n")
        f.write("%-96s %6s %s %4s %7s %5s\n" % ("kernel", "instr", " ".join("%8s" % c for c in cols), "regs", "smem_B", "local"))
        for name, c, reg, sh, loc in rows:
            f.write("%-96s %6d %s %4s %7s %5s\n" % (name, c["instr"], " ".join("%8d" % c[k] for k in cols), reg, sh, loc))
        tot = dict((k, sum(r[1][k] for r in rows)) for k in cols)
        f.write("%-96s %6d %s\n" % ("TOTAL", sum(r[1]["instr"] for r in rows), " ".join("%8d" % tot[k] for k in cols)))
    print("wrote", dst, "kernels:", len(rows))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_sass_evidence.txt"))
