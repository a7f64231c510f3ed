"""Device-resident rate of the any-alignment kernels on sizes that are not multiples of 16/32 (developer tool)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, jmcodec_b200 as J
ctx = J.Ctx(0)
peak, _ = bench.peaks()
if len(sys.argv) > 2 and sys.argv[1] == "--spec":          # --spec op,w,h,pitch,frames [more specs ...]
    for sp in sys.argv[2:]:
        op, w, h, pitch, n = sp.split(",")
        bench.WORKLOADS["_x"] = (op, int(w), int(h), int(pitch), int(n))
        r = bench.device_only(ctx, "_x", 0, 10, 3)
        print(f"{sp:28s} {r['frames_per_s']:12.0f} fps {r['gbs']:8.1f} GB/s  {r['gbs']/peak:.3f} of peak")
    sys.exit(0)
for name, spec in {
    "1080x1920 portrait p1088": ("i420", 1080, 1920, 1088, 150), "1080x1920 portrait p1280": ("i420", 1080, 1920, 1280, 150),
    "1366x768 p1536": ("i420", 1366, 768, 1536, 300), "854x480 p1024": ("i420", 854, 480, 1024, 600),
    "1918x1078 p2048": ("i420", 1918, 1078, 2048, 300), "1919x1079 p2048": ("i420", 1919, 1079, 2048, 300),
    "720x480 p768": ("i420", 720, 480, 768, 800), "1280x720 p1280": ("i420", 1280, 720, 1280, 600),
    "pack 1080x1920 p1088": ("pack", 1080, 1920, 1088, 150), "pack 1366x768 p1536": ("pack", 1366, 768, 1536, 300),
    "pack 854x480 p1024": ("pack", 854, 480, 1024, 600), "pack 1919x1079 p2048": ("pack", 1919, 1079, 2048, 300),
    "nv12 1366x768 p1536": ("nv12", 1366, 768, 1536, 300), "rgb 1080x1920 p1088": ("rgb", 1080, 1920, 1088, 150),
    "rgb 1366x768 p1536": ("rgb", 1366, 768, 1536, 300), "argb 1366x768 p1536": ("argb", 1366, 768, 1536, 300),
    "fused 1366x768 p1536": ("fused", 1366, 768, 1536, 300),
    "rgb 1376x768 p1536": ("rgb", 1376, 768, 1536, 300), "rgb 1536x768 p1536": ("rgb", 1536, 768, 1536, 300),
    "argb 1376x768 p1536": ("argb", 1376, 768, 1536, 300), "argb 1536x768 p1536": ("argb", 1536, 768, 1536, 300),
    "rgb2nv12 1376x768 p1536": ("rgb2nv12", 1376, 768, 1536, 300), "rgb2nv12 1536x768 p1536": ("rgb2nv12", 1536, 768, 1536, 300), "argb 3840x2160 p4096": ("argb", 3840, 2160, 4096, 64),
    "argb 1080x1920 p1088": ("argb", 1080, 1920, 1088, 150), "rgb 1280x720 p1280": ("rgb", 1280, 720, 1280, 400),
    "rgb 3840x2160 p4096": ("rgb", 3840, 2160, 4096, 64), "fused 3840x2160 p4096": ("fused", 3840, 2160, 4096, 64),
    "rgb 1920x1080 p2048": ("rgb", 1920, 1080, 2048, 200), "fused 1920x1080 p2048": ("fused", 1920, 1080, 2048, 200),
    "rgb2nv12 3840x2160 p4096": ("rgb2nv12", 3840, 2160, 4096, 64), "rgb2nv12 1920x1080 p2048": ("rgb2nv12", 1920, 1080, 2048, 200),
    "rgb2nv12 1366x768 p1536": ("rgb2nv12", 1366, 768, 1536, 300),
}.items():
    if len(sys.argv) > 1 and not any((a[1:] == name) if a.startswith('=') else (a in name) for a in sys.argv[1:]):
        continue
    bench.WORKLOADS["_x"] = spec
    r = bench.device_only(ctx, "_x", 0, 10, 3)
    print(f"{name:28s} {r['frames_per_s']:12.0f} fps {r['gbs']:8.1f} GB/s  {r['gbs']/peak:.3f} of peak")
