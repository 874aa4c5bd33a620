"""Development probe: the Delta look-back path through the library (registry C ABI), on the probe's workload
(one segment, random deltas), for a few grid limits; prints the resolved launch configuration and the time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import libgiddy_b200 as g
from libgiddy_b200 import schemes as S

lib = g.default_library()
dev = torch.device("cuda")
for (U, C, n, L) in [(8, 4, 1 << 26, 1 << 26), (4, 1, 1 << 27, 1 << 27), (8, 4, 1 << 26, 1 << 20), (8, 4, (1 << 26), (1 << 26) - 64)]:
    cin = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev) if C == 1 else torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device=dev)
    num_segments = (n + L - 1) // L
    anchors = torch.zeros(num_segments, dtype=torch.int64 if U == 8 else torch.int32, device=dev)
    out = torch.zeros(n + 64, dtype=torch.int64 if U == 8 else torch.int32, device=dev)
    for limits in (None, {"grid_size": 444}, {"grid_size": 296}, {"grid_size": 148}):
        cfg = S.delta(out, cin, anchors, n, L, uncompressed_size=U, compressed_size=C, library=lib, limits=limits)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            S.delta(out, cin, anchors, n, L, uncompressed_size=U, compressed_size=C, library=lib, limits=limits)
        e0.record()
        for _ in range(20):
            S.delta(out, cin, anchors, n, L, uncompressed_size=U, compressed_size=C, library=lib, limits=limits)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("U=%d C=%d n=%d L=%d limits=%s config=%s: %.3f ms  %.0f GB/s" % (U, C, n, L, limits, cfg, ms, n * (U + C) / ms / 1e6), flush=True)
