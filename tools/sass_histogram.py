#!/usr/bin/env python
"""Per-kernel SASS histogram of the data-movement instructions (cuobjdump -sass), for profiles/.

    python tools/sass_histogram.py jmcodec_b200/libjmcodec_b200.so tools/tma_ab > profiles/r2_sass_histogram.txt

UBLKCP = cp.async.bulk (1-D bulk copy engine), UTMALDG / UTMASTG = cp.async.bulk.tensor (tensor-map TMA),
SYNCS = mbarrier operations, LDG / STG = register-staged global loads / stores, LDS / STS = shared memory,
PRMT = byte permute (U/V de-/interleave), IDP = dp2a/dp4a (colour arithmetic)."""
import collections
import re
import subprocess
import sys

OPS = ["UBLKCP", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "LDG", "STG", "LDS", "STS", "PRMT", "IDP", "SHF", "BAR", "ACQBULK", "UTMACMDFLUSH"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    for path in sys.argv[1:]:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        kernels, cur = collections.OrderedDict(), None
        for ln in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                cur = m.group(1)
                kernels[cur] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m and cur:
                op = m.group(1)
                kernels[cur]["_total"] += 1
                for o in OPS:
                    if op == o or op.startswith(o + "."):
                        kernels[cur][o] += 1
        names = demangle(list(kernels))
        print(f"== {path}")
        print(f"{'kernel':<86} {'instr':>6} " + " ".join(f"{o:>7}" for o in OPS[:11]))
        for k, c in kernels.items():
            n = re.sub(r"\(.*", "", names.get(k, k)).replace("void ", "").replace("jmc::", "")
            print(f"{n[:86]:<86} {c['_total']:>6} " + " ".join(f"{c[o]:>7}" for o in OPS[:11]))
        print()


if __name__ == "__main__":
    main()
