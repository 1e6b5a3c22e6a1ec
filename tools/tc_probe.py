"""Probe the tcgen05 conv kernels against the CUDA-core engine on the GPU (no oracle, fast).
usage: python tools/tc_probe.py   (spawns one subprocess per descriptor variant, each under a timeout)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from multi_stylegan_b200 import _C, _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    print("tc available:", _C.tensor_core_path_available(), flush=True)
    shapes = [  # B, C, O, H, W, k, s, p, per_sample
        (2, 64, 64, 32, 32, 3, 1, 1, True),
        (2, 64, 64, 32, 32, 1, 1, 0, False),
        (2, 32, 48, 16, 16, 3, 1, 1, True),
        (1, 96, 272, 64, 64, 3, 1, 1, False),
        (2, 48, 40, 32, 32, 2, 2, 0, True),
        (2, 32, 32, 64, 64, 3, 2, 0, False),
    ]
    which = os.environ.get("PROBE_WHICH", "fdw")
    for shp in shapes:
        B, C, O, H, W, k, s, p, per = shp
        x = torch.randn(B, C, H, W, device=dev)
        w = torch.randn((B, O, C, k, k) if per else (O, C, k, k), device=dev) / (C * k * k) ** 0.5
        res = {}
        for name, flags in (("simt", _lib.CONV_FORCE_SIMT), ("tc", _lib.CONV_FORCE_TC)):
            _C.conv_flags = flags
            try:
                y = _C.conv2d_forward(x, w, s, p) if "f" in which else None
                if name == "simt" or y is None:
                    dy = torch.randn(B, O, (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1, device=dev) \
                        if "dy" not in res else res["dy"]
                    res["dy"] = dy
                dy = res["dy"]
                dx = _C.conv2d_dgrad(dy, w, (H, W), s, p) if "d" in which else None
                dw = _C.conv2d_wgrad(dy, x, (k, k), s, p, per) if "w" in which else None
                torch.cuda.synchronize()
                res[name] = (y, dx, dw)
            except RuntimeError as e:
                print(shp, name, "ERROR", str(e)[:200], flush=True)
                res[name] = None
            if name == "tc" and os.environ.get("MSG_B200_TC_DEBUG") == "1":
                import ctypes
                import numpy as np
                words = ctypes.c_size_t(0)
                ptr = _lib.lib().msg_debug_buffer(ctypes.byref(words))
                if ptr:
                    print("   markers:", [hex(int(v)) for v in np.ctypeslib.as_array(ptr, shape=(8,))], flush=True)
        if res.get("simt") and res.get("tc"):
            errs = []
            for a, b in zip(res["tc"], res["simt"]):
                errs.append(None if a is None else round(((a - b).abs().max() / b.abs().max()).item(), 5))
            print(shp, "rel err fwd/dgrad/wgrad:", errs, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    variants = sys.argv[1:] or ["0", "1", "2", "3"]
    for v in variants:
        for which in ("f", "d", "w"):
            dbg = "1" if v.endswith("d") else "0"
            env = dict(os.environ, MSG_B200_TC_VARIANT=v.rstrip("d"), PROBE_WHICH=which, MSG_B200_TC_DEBUG=dbg)
            print("=== variant", v, "which", which, flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, "child"], env=env, timeout=120, capture_output=True, text=True)
                print(r.stdout[-3000:], r.stderr[-1500:], "exit", r.returncode, flush=True)
            except subprocess.TimeoutExpired:
                print("TIMEOUT", flush=True)
