"""Row pairs per CTA of the re-aligning bulk colour kernel (JMC_RGB_BULK_PAIRS), on frames whose rows are not 16-byte
multiples (developer tool).  Buffers are built once per geometry; the switch is re-read between launches
(jmc_reload_env).  pairs=0 is what the library picks on its own."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, jmcodec_b200 as J

ctx = J.Ctx(0)
peak, _ = bench.peaks()
SPECS = [("rgb", 1366, 768, 1536, 300), ("fused", 1366, 768, 1536, 300), ("rgb", 1080, 1920, 1088, 150),
         ("fused", 1080, 1920, 1088, 150), ("rgb", 854, 480, 1024, 600), ("fused", 854, 480, 1024, 600),
         ("rgb", 1920, 1080, 2048, 200), ("fused", 1920, 1080, 2048, 200)]
for op, w, h, pitch, n in SPECS:
    in_b, out_b, out2_b = bench.io_bytes(op, w, h, pitch)
    host = np.concatenate(bench.make_inputs(op, w, h, pitch, 0)[:min(n, bench.N_DISTINCT)])
    d_in = ctx.alloc(in_b * n)
    for r in range(-(-n // bench.N_DISTINCT)):
        cnt = min(bench.N_DISTINCT, n - r * bench.N_DISTINCT)
        ctx.h2d(d_in + r * bench.N_DISTINCT * in_b, host, cnt * in_b)
    d_out = ctx.alloc(out_b * n)
    d_out2 = ctx.alloc(out2_b * n) if out2_b else None
    j = bench.build_job(ctx, op, w, h, pitch, n, d_in, d_out, d_out2)
    alg = ctx.algorithmic_bytes(j) * n
    out = []
    for pairs in (0, 1, 2, 3, 4, 6):
        os.environ["JMC_RGB_BULK_PAIRS"] = str(pairs)
        J.reload_env()
        for _ in range(3):
            ctx.convert(j)
        ctx.sync()
        ms = ctx.convert_timed(j, 10)
        out.append(f"pairs {pairs}: {alg / (ms * 1e-3) / 1e9 / peak:.3f}")
    os.environ.pop("JMC_RGB_BULK_PAIRS", None)
    J.reload_env()
    for d in (d_in, d_out, d_out2):
        if d:
            ctx.free(d)
    print(f"{op:6s} {w}x{h} pitch {pitch} x{n}:  " + "  ".join(out))
