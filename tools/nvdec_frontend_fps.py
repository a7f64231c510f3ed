"""Frames/s of the NVDEC front-end of jm_nvdec_* against tests/fake_nvcuvid (no NVDEC engine is exposed on this pool):
batch drain (up to 8 displayed pictures mapped and converted per launch, unmapped on the convert event) vs the
reference's scheme (one picture mapped, converted and unmapped per call: map_limit 1).  Developer tool.

The fake's "decode" is a synchronous host->device copy of the picture, so absolute numbers say little; the difference
between the two settings is the front-end's own per-frame cost (map / launch / event / unmap)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["JMC_NVCUVID_LIB"] = os.path.join(ROOT, "tests", "fake_nvcuvid", "libfake_nvcuvid.so")
import fake_stream as FS          # noqa: E402
import jmcodec_b200 as J          # noqa: E402
from jmcodec_b200 import synth    # noqa: E402


def run(w, h, n, per_packet, map_limit, delay):
    need = w * h * 3 // 2
    pics = [FS.picture(synth.random_bytes(need, synth.frame_key(70, f))) for f in range(8)]
    packets = [np.concatenate([pics[(i + k) % 8] for k in range(per_packet)]) for i in range(0, n, per_packet)]
    dec = J.NvDec(0)
    dec.set_option("map_limit", map_limit)
    dec.set_display_delay(delay)
    assert dec.init(0, 1) == 0, J.last_error()
    out = dec.alloc_host(need)
    dec.decode_frame(FS.sequence_header(w, h))
    got = 0
    t0 = time.perf_counter()
    for p in packets:
        r, g = dec.decode_frame(p)
        if g == 1:
            dec.output_frame(out, need)
            got += 1
        # a packet with several pictures leaves frames queued: drain them the way a player would, one call each
        for _ in range(per_packet - 1):
            if dec.is_exit():
                break
            # no new data: an empty-but-valid call is not part of the reference API, so feed a filler NAL the fake ignores
            r, g = dec.decode_frame(np.array([0, 0, 1, 0x09, 0x80], np.uint8))
            if g == 1:
                dec.output_frame(out, need)
                got += 1
    while not dec.is_exit():
        r, g = dec.decode_frame(None, 0)
        if g == 1:
            dec.output_frame(out, need)
            got += 1
    dt = time.perf_counter() - t0
    launches = dec.launches
    dec.free_host(out)
    dec.deinit()
    return {"frames": got, "fps": round(got / dt, 1), "launches": launches}


if __name__ == "__main__":
    res = {}
    for (w, h, n) in ((1920, 1080, 256), (640, 360, 512)):
        for per_packet in (1, 8):
            for map_limit in (8, 1):
                for delay in (0, 2):
                    key = f"{w}x{h} pictures_per_packet={per_packet} map_limit={map_limit} display_delay={delay}"
                    res[key] = run(w, h, n, per_packet, map_limit, delay)
                    print(key, res[key], flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_nvdec_frontend_fps.json"), "w"), indent=1)
