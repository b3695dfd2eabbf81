"""Device-resident timing of the open-loop generator kernels (tg_openloop_type1 / type2).

    python tools/bench_openloop.py [--batch 65536] [--T 1200] [--reps 5] [--kinds type1 type2] [--no-modes]

Prints one JSON line per kind: trajectory-steps/s, kernel ms, achieved output bandwidth (algorithmic bytes =
112 B per trajectory-step: 6 clean + 6 noisy + 2 control doubles, + 1 B of mode for type 2) against the measured
HBM peak of MEASURED_PEAKS.json.  Timed with CUDA events on the handle's stream after 2 warm-up launches; a 256 MB
buffer is zeroed between launches (L2 flush)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import trajectory_generation_b200 as tg  # noqa: E402
from trajectory_generation_b200 import _lib  # noqa: E402


def hbm_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs", "hbm_gbps"):
            if k in pk:
                return float(pk[k]), f"MEASURED_PEAKS.json:{k}"
        for k, v in pk.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), f"MEASURED_PEAKS.json:{k}"
    except Exception:
        pass
    return 6540.5, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--T", type=int, default=1200)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--kinds", nargs="+", default=["type1", "type2"])
    ap.add_argument("--no-modes", action="store_true")
    a = ap.parse_args()
    B, T = a.batch, a.T
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak, src = hbm_peak()
    for kind in a.kinds:
        gen = tg.OpenLoopGenerator(kind, Ts=0.01)
        gen.set_stream(stream.cuda_stream)
        x0 = gen.sample_x0(min(B, 4096), seed=42)
        x0 = np.ascontiguousarray(np.tile(x0, ((B + len(x0) - 1) // len(x0), 1))[:B])
        d_x0 = _lib.DeviceBuffer(B * 48)
        _lib.check(_lib.load().tg_memcpy_h2d(gen._h, d_x0.ptr, x0.ctypes.data, B * 48))
        d_cl, d_no, d_U = _lib.DeviceBuffer(B * (T + 1) * 48), _lib.DeviceBuffer(B * (T + 1) * 48), _lib.DeviceBuffer(B * T * 16)
        n_modes = B if kind == "type1" else B * T
        d_m = None if a.no_modes else _lib.DeviceBuffer(n_modes)
        ms = []
        with torch.cuda.stream(stream):
            for rep in range(a.reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                gen.generate_device(d_x0.ptr, B, T, d_cl.ptr, d_no.ptr, d_U.ptr, d_m.ptr if d_m else None)
                e1.record(stream)
                e1.synchronize()
                if rep >= 2:
                    ms.append(e0.elapsed_time(e1))
        ms_avg = float(np.mean(ms))
        steps = B * T
        bytes_alg = B * (T + 1) * 96 + B * T * 16 + (0 if a.no_modes else n_modes) + B * 48
        gbps = bytes_alg / (ms_avg * 1e-3) / 1e9
        print(json.dumps({"kernel": f"tg_openloop_kernel<{kind}>", "batch": B, "T": T, "ms": ms_avg, "ms_min": float(np.min(ms)),
                          "traj_steps_per_s": steps / (ms_avg * 1e-3), "algorithmic_bytes": bytes_alg,
                          "roofline": {"bound": "hbm", "achieved": gbps, "peak": peak, "unit": "GB/s", "frac": gbps / peak,
                                       "peak_source": src}}))
        for buf in (d_x0, d_cl, d_no, d_U, d_m):
            if buf:
                buf.free()
        gen.close()


if __name__ == "__main__":
    main()
