"""CUDA-event timing of the cube-boundary kernels (csrc/cubeio.cu) at the C3 slab size: Q/U (1000, 512, 512) ->
16384-pixel planar slabs, planar Faraday spectra -> (2, 7968, Y, X) planes, per-channel nanstd.  One JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csromer_b200 as cs  # noqa: E402

m, Y, X, P, n = 1000, 512, 512, 16384, 7968
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
Q = torch.randn((m, Y, X), generator=g, device=dev)
U = torch.randn((m, Y, X), generator=g, device=dev)
cube = cs.DeviceCube(Q, U)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
ms = timed(lambda i=0: cube.slab((i * P) % (Y * X - P), P))
nbytes = 2 * m * P * 4 * 2 + m * P          # read Q, U; write planar + mask (torch.zeros of the output not counted)
res["pack_slab"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"],
                    "note": "includes the torch.zeros of the output slab (another write pass)"}
F = torch.randn((P, 2 * n), generator=g, device=dev)
out = torch.zeros((2, n, Y // 4, X), device=dev)   # a quarter image keeps the buffer at 8 GB
small = cs.DeviceCube(Q[:, :Y // 4], U[:, :Y // 4])
ms = timed(lambda i=0: small.store(F, n, (i * P) % (Y // 4 * X - P), out))
nbytes = 2 * n * P * 4 * 2
res["unpack_slab"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"]}
ms = timed(lambda i=0: cube.channel_stats("Q"), reps=5)
nbytes = m * Y * X * 4                              # the second (mean-centred) pass re-reads the plane: mostly L2 hits
res["channel_stats"] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm_gbs"]}
print(json.dumps({"workload": "cube boundary kernels, Q/U (1000, 512, 512) fp32, 16384-pixel slabs, n_phi=7968", "kernels": res}))
