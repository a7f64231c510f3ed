"""Shared case matrix for the parity tests and the golden-vector generator.

Every case is a plain dict so that its id is stable and the golden JSON can be keyed by it.
"""
from __future__ import annotations

import hashlib

import numpy as np

from jmcodec_b200 import synth

SLACK = 64          # sentinel bytes after the reference's w*h*3/2 output, must stay untouched

# (w, h, pitch): full sizes, non-multiple-of-32 widths, odd dimensions, degenerate sizes
NVDEC_SIZES = [
    (1920, 1080, 2048), (1920, 1080, 1920), (3840, 2160, 4096), (3840, 2160, 3840),
    (1918, 1078, 2048), (1919, 1079, 2048), (1280, 720, 1280), (720, 480, 768),
    (176, 144, 256), (64, 64, 64), (48, 32, 64), (40, 24, 48), (36, 20, 40), (34, 18, 36),
    (16, 16, 256), (6, 6, 8), (5, 3, 8), (3, 5, 4), (2, 2, 256), (1, 1, 256), (1, 2, 1), (2, 1, 2),
]
SMALL = 1 << 20     # cases whose surface is below this are cheap enough for every CPU run


def nvdec_cases():
    out = []
    for (w, h, p) in NVDEC_SIZES:
        for fmt in (0, 1):
            kinds = ("random", "gradient") if w * h <= 1920 * 1080 else ("random",)
            for kind in kinds:
                out.append(dict(op="nvdec", w=w, h=h, pitch=p, fmt=fmt, kind=kind))
    out.append(dict(op="nvdec", w=64, h=64, pitch=64, fmt=7, kind="random"))    # any non-zero fmt is I420
    return out


# (pitch, surf_rows, crop_x, crop_y, crop_w, crop_h); UV plane starts at row `surf_rows`
INTEL_GEOMS = [
    (1920, 1088, 0, 0, 1920, 1080), (1952, 1088, 16, 8, 1920, 1072), (1952, 1088, 6, 4, 1280, 720),
    (1952, 1088, 5, 3, 1279, 719), (3840, 2176, 0, 0, 3840, 2160), (64, 32, 0, 0, 48, 32),
    (64, 32, 2, 2, 33, 17), (64, 32, 3, 1, 16, 16), (32, 32, 0, 0, 2, 2), (32, 32, 1, 1, 1, 1),
]


def inteldec_cases():
    return [dict(op="inteldec", pitch=p, rows=r, cx=cx, cy=cy, cw=cw, ch=ch, fmt=fmt)
            for (p, r, cx, cy, cw, ch) in INTEL_GEOMS for fmt in (0, 1)]


def intelenc_cases():
    out = [dict(op="intelenc", pitch=p, rows=r, cx=cx, cy=cy, cw=cw, ch=ch, i420=i)
           for (p, r, cx, cy, cw, ch) in INTEL_GEOMS for i in (0, 1)]
    # CropW/CropH == 0 -> the reference falls back to Info.Width/Height (intel_enc.cpp:271-278)
    out.append(dict(op="intelenc", pitch=64, rows=32, cx=0, cy=0, cw=0, ch=0, i420=1, info_w=48, info_h=32))
    out.append(dict(op="intelenc", pitch=64, rows=32, cx=0, cy=0, cw=0, ch=0, i420=0, info_w=48, info_h=32))
    return out


NVENC_SIZES = [(1920, 1080, 2048), (3840, 2160, 4096), (1280, 720, 1536), (1919, 1079, 2048),
               (1918, 1078, 2048), (64, 64, 512), (34, 18, 512), (6, 6, 512), (5, 3, 512), (2, 2, 512)]
NVENC_FMTS = {"nv12": 0x1, "yv12": 0x10, "argb": 0x01000000, "abgr": 0x10000000}


def nvenc_cases():
    out = []
    for (w, h, s) in NVENC_SIZES:
        for name in ("nv12", "yv12"):
            out.append(dict(op="nvenc", w=w, h=h, stride=s, fmt=name))
    for (w, h, s) in [(64, 64, 512), (34, 18, 512), (5, 3, 512), (1280, 720, 5120)]:
        out.append(dict(op="nvenc", w=w, h=h, stride=max(s, w * 4), fmt="argb"))
    return out


RGB_SIZES = [(1920, 1080, 2048), (3840, 2160, 4096), (1918, 1078, 2048), (1919, 1079, 2048),
             (1280, 720, 1280), (64, 64, 64), (48, 32, 64), (34, 18, 36), (6, 6, 8), (5, 3, 8), (3, 5, 4), (2, 2, 256)]


def rgb_cases():
    out = []
    for (w, h, p) in RGB_SIZES:
        kinds = ["random"]
        if w * h <= 64 * 64:
            kinds += ["gradient", "const0", "const16", "const128", "const235", "const240", "const255"]
        for k in kinds:
            out.append(dict(op="rgb24", w=w, h=h, pitch=p, kind=k))
    return out


def argb_cases():
    out = []
    for (w, h, p) in RGB_SIZES:
        kinds = ["random"] + (["gradient", "const16", "const235", "const255"] if w * h <= 64 * 64 else [])
        for k in kinds:
            out.append(dict(op="argb32", w=w, h=h, pitch=p, kind=k))
    return out


def rgb2nv12_cases():
    """RGB24 -> pitched NV12 (builder-defined forward BT.601): rgb_pitch = 3*w + skew, surface pitch p."""
    out = []
    for (w, h, p) in RGB_SIZES + [(1366, 768, 1536), (1, 1, 16), (2, 1, 16), (1, 2, 16)]:
        for skew in ((0, 5) if w * h <= 64 * 64 else (0,)):
            kinds = ["random"] + (["const0", "const255", "primaries"] if w * h <= 64 * 64 and skew == 0 else [])
            for k in kinds:
                out.append(dict(op="rgb2nv12", w=w, h=h, pitch=p, skew=skew, kind=k))
    return out


def all_cases():
    return nvdec_cases() + inteldec_cases() + intelenc_cases() + nvenc_cases() + rgb_cases() + argb_cases() + rgb2nv12_cases()


def case_id(c: dict) -> str:
    return "-".join(f"{k}={c[k]}" for k in sorted(c))


def is_small(c: dict) -> bool:
    if "pitch" in c and "h" in c:
        return c["pitch"] * c["h"] <= SMALL
    if "rows" in c:
        return c["pitch"] * c["rows"] <= SMALL
    return c["stride"] * c["h"] <= SMALL


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------------------------------------
# Input builders + checker runners.  `chk` is an oracle.Checker (port or compiled reference).
# Each runner returns (ret, out_len, output_bytes) where output_bytes includes sentinel regions.
# ---------------------------------------------------------------------------------------------
def nvdec_input(c):
    return synth.nv12_surface(c["w"], c["h"], c["pitch"], stream=1, frame=c["w"] * 7 + c["h"], kind=c["kind"])


def run_nvdec(chk, c):
    s = nvdec_input(c)
    cap = c["w"] * c["h"] * 3 // 2 + SLACK
    out = np.full(cap, synth.OUT_FILL, np.uint8)
    r, n = chk.nvdec_output_frame(s, c["pitch"], c["w"], c["h"], c["fmt"], out, cap)
    return r, n, out


def intel_surface(c):
    """Full-height NV12 system-memory surface: `rows` luma rows then rows/2 chroma rows."""
    rows = c["rows"]
    return synth.nv12_surface(c["pitch"], rows, c["pitch"], stream=2, frame=c["cw"] * 3 + c["cx"],
                              rows=rows + rows // 2)


def run_inteldec(chk, c):
    s = intel_surface(c)
    cap = c["cw"] * c["ch"] + (c["cw"] * c["ch"]) // 2 + SLACK
    out = np.full(cap, synth.OUT_FILL, np.uint8)
    r, n = chk.inteldec_output_frame(s, c["pitch"] * c["rows"], c["pitch"],
                                     (c["cx"], c["cy"], c["cw"], c["ch"]), c["fmt"], out, cap)
    return r, n, out


def intelenc_dims(c):
    if c["cw"] > 0 and c["ch"] > 0:
        return c["cw"], c["ch"]
    return c["info_w"], c["info_h"]


def intelenc_input(c):
    w, h = intelenc_dims(c)
    return synth.i420_frame(w, h, stream=3, frame=w + h + c["i420"])


def run_intelenc(chk, c):
    yuv = intelenc_input(c)
    rows = c["rows"]
    surf = np.full(c["pitch"] * (rows + rows // 2), synth.PAD_BYTE, np.uint8)
    r = chk.intelenc_input(yuv, c["i420"], surf, c["pitch"] * rows, c["pitch"],
                           (c.get("info_w", c["pitch"]), c.get("info_h", rows)),
                           (c["cx"], c["cy"], c["cw"], c["ch"]))
    return r, surf.size, surf


def nvenc_input(c):
    w, h = c["w"], c["h"]
    n = w * h * 4 if c["fmt"] in ("argb", "abgr") else w * h * 3 // 2 + 8
    return synth.random_bytes(n, synth.frame_key(4, w * 5 + h))


def nvenc_surface_bytes(c):
    w, h, s = c["w"], c["h"], c["stride"]
    return s * h if c["fmt"] in ("argb", "abgr") else s * (h * 3 // 2 + 1)


def run_nvenc(chk, c):
    import oracle
    surf = np.full(nvenc_surface_bytes(c), synth.PAD_BYTE, np.uint8)
    if chk.kind == "reference":       # the reference's own host code over a fake CUDA driver
        r, _ = oracle.ref_nvenc_convert(nvenc_input(c), NVENC_FMTS[c["fmt"]], c["w"], c["h"], surf, c["stride"])
    else:
        r = oracle.nvenc_upload(nvenc_input(c), NVENC_FMTS[c["fmt"]], c["w"], c["h"], surf, c["stride"])
    return r, surf.size, surf


def rgb_input(c):
    return synth.nv12_surface(c["w"], c["h"], c["pitch"], stream=5, frame=c["w"] + c["h"], kind=c["kind"])


def run_rgb(_chk, c):
    import oracle
    rgb_pitch = c["w"] * 3
    out = np.full(rgb_pitch * c["h"] + SLACK, synth.OUT_FILL, np.uint8)
    r = oracle.nv12_to_rgb24(rgb_input(c), c["pitch"], c["w"], c["h"], out, rgb_pitch)
    return r, rgb_pitch * c["h"], out


def run_argb(_chk, c):
    import oracle
    pitch4 = c["w"] * 4
    out = np.full(pitch4 * c["h"] + SLACK, synth.OUT_FILL, np.uint8)
    r = oracle.nv12_to_argb32(rgb_input(c), c["pitch"], c["w"], c["h"], out, pitch4)
    return r, pitch4 * c["h"], out


def rgb2nv12_input(c):
    """Tight-ish RGB24 rows: 3*w bytes of pixels then `skew` pad bytes per row."""
    w, h, rp = c["w"], c["h"], 3 * c["w"] + c["skew"]
    if c["kind"] == "random":
        a = synth.random_bytes(rp * h, synth.frame_key(9, w * 131 + h) ^ 0x0F0F0F0F)
    elif c["kind"] == "primaries":
        pal = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255], [255, 0, 255], [255, 255, 255], [0, 0, 0]], np.uint8)
        a = np.full((h, rp), synth.PAD_BYTE, np.uint8)
        for y in range(h):
            a[y, :3 * w] = pal[(np.arange(w) // 2 + y // 2) % 8].reshape(-1)
        a = a.reshape(-1)
    else:
        a = np.full(rp * h, int(c["kind"][5:]), np.uint8)
    return a


def rgb2nv12_surface_bytes(c):
    return c["pitch"] * (c["h"] + (c["h"] >> 1) + 1)


def run_rgb2nv12(_chk, c):
    import oracle
    surf = np.full(rgb2nv12_surface_bytes(c), synth.PAD_BYTE, np.uint8)
    r = oracle.rgb24_to_nv12(rgb2nv12_input(c), 3 * c["w"] + c["skew"], c["w"], c["h"], surf, c["pitch"])
    return r, surf.size, surf


RUNNERS = {"rgb2nv12": run_rgb2nv12, "argb32": run_argb, "nvdec": run_nvdec, "inteldec": run_inteldec, "intelenc": run_intelenc,
           "nvenc": run_nvenc, "rgb24": run_rgb}
# ops for which the unmodified reference has CPU code that oracle/_ref executes
REF_OPS = ("nvdec", "inteldec", "intelenc", "nvenc")


def run_case(chk, c):
    return RUNNERS[c["op"]](chk, c)
