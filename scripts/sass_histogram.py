#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libfrustum_b200.so (cuobjdump -sass): proves which kernels carry tcgen05
MMA (UTCHMMA / UTCQMMA), TMEM loads (LDTM), TMA tensor loads / stores (UTMALDG / UTMASTG), bulk copies (UBLKCP),
cluster barriers.  Writes profiles/<tag>_sass_histogram.txt.  Runs on the CPU (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "frustum_convnet_b200", "libfrustum_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF",
        "SYNCS", "ELECT", "FFMA", "HMMA", "ATOMG", "REDG", "RED", "MEMBAR", "ERRBAR", "CCTL"]


def main(tag):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern = None
    hist = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kern = re.sub(r"\(.*", "", kern).replace("fcn::", "")
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and kern:
            op = m.group(1)
            full = op + (m.group(2) or "")
            hist[kern][op] += 1
            if op in ("UTCHMMA", "UTCBAR") and ".2CTA" in full:
                hist[kern][op + ".2CTA"] += 1
    lines = ["SASS opcode histogram of libfrustum_b200.so (cuobjdump -sass, sm_100a) - %s" % tag,
             "UTCHMMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld (TMEM), UTMALDG/UTMASTG = TMA tensor load/store,",
             "UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, FFMA = fp32 FMA (SIMT kernels)", ""]
    cols = KEYS + ["UTCHMMA.2CTA"]
    lines.append("%-58s %7s " % ("kernel", "instrs") + " ".join("%8s" % c[:8] for c in cols))
    for k, h in hist.items():
        tot = sum(v for kk, v in h.items() if "." not in kk)
        lines.append("%-58s %7d " % (k[:58], tot) + " ".join("%8d" % h.get(c, 0) for c in cols))
    path = os.path.join(ROOT, "profiles", "%s_sass_histogram.txt" % tag)
    open(path, "w").write("\n".join(lines) + "\n")
    print(path)
    for line in lines[4:]:
        if any(x in line for x in ("pointnet_tc", "mega", "conv_gemm_t", "kernel ")):
            print(line[:200])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
