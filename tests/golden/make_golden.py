#!/usr/bin/env python
"""Generate tests/golden/kat_sha256.json -- known-answer vectors for the surface-format path.

The reference repository has no golden vectors of its own, so these are made
by EXECUTING THE UNMODIFIED REFERENCE CODE, compiled in place from /root/reference into
oracle/_ref/libjmref.so (oracle/Makefile), on the deterministic synthetic surfaces of
jmcodec_b200/synth.py.  Each entry records the reference's return code, *out_len and the SHA-256
of the whole output buffer INCLUDING its 0xA5 / 0xCD pre-filled slack and padding, so bytes the
reference leaves untouched are pinned too.  Small cases also carry the raw bytes (hex).

The nv_enc entries come from the reference's own nvenc_convert_yuv_data_to_nv12() executed against a
fake CUDA driver (oracle/ref_nvenc_driver.cpp; only the absent InterleaveUV PTX is emulated).
Entries with "source": "port" have no executable reference code (RGB24: no YUV->RGB in the reference
at all); they come from the C restatement oracle/jm_oracle.c and are marked "parity": "unpinned".

Run in the dev container (needs /root/reference or a prebuilt oracle/_ref):
    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import cases as K  # noqa: E402


def main():
    oracle.build()
    ref = oracle.ref()
    port = oracle.port()
    table = {}
    for c in K.all_cases():
        from_ref = c["op"] in K.REF_OPS
        r, n, out = K.run_case(ref if from_ref else port, c)
        e = {"ret": int(r), "out_len": int(n), "nbytes": int(out.size), "sha256": K.sha(out),
             "source": "reference" if from_ref else "port"}
        if c["op"] in ("rgb24", "argb32", "rgb2nv12"):
            e["parity"] = "unpinned"
        if out.size <= 512:
            e["hex"] = out.tobytes().hex()
        table[K.case_id(c)] = e
    path = os.path.join(ROOT, "tests", "golden", "kat_sha256.json")
    with open(path, "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
        f.write("\n")
    print(f"wrote {len(table)} vectors to {path}")


if __name__ == "__main__":
    main()
