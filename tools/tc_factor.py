"""One-factor-at-a-time probe of the tcgen05 forward/wgrad kernels (debug mode, subprocess per case)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {  # name: (B, C, O, H, W, k, pad, per_sample)
    "3x3_same": (2, 64, 64, 32, 32, 3, 1, True),
    "3x3_same_16": (2, 64, 64, 16, 16, 3, 1, True),
    "3x3_same_4": (2, 64, 64, 4, 4, 3, 1, True),
    "3x3_same_128_bigN": (1, 96, 272, 128, 128, 3, 1, False),
    "1x1_N3": (2, 64, 3, 32, 32, 1, 0, True),
    "3x3_C6": (2, 6, 32, 64, 64, 3, 1, False),
    "3x1_valid_Hshift": (1, 32, 64, 34, 32, (3, 1), 0, False),
    "1x3_valid_Wshift": (1, 32, 64, 32, 34, (1, 3), 0, False),
    "3x1_same_negH": (1, 32, 64, 32, 32, (3, 1), (1, 0), False),
    "1x1_shared": (1, 32, 64, 32, 32, 1, 0, False),
    "1x1_persample": (2, 32, 64, 32, 32, 1, 0, True),
    "1x1_bigger_K": (1, 256, 64, 32, 32, 1, 0, False),
    "1x1_40x40_posOOB": (1, 32, 64, 40, 40, 1, 0, False),
    "3x3_valid": (1, 32, 64, 34, 34, 3, 0, False),
    "3x3_same_negOOB": (1, 32, 64, 32, 32, 3, 1, False),
}

def child(name, which):
    import numpy as np, torch
    from multi_stylegan_b200 import _C, _lib
    B, C, O, H, W, k, p, per = CASES[name]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device=dev)
    kh, kw = (k, k) if isinstance(k, int) else k
    w = torch.randn((B, O, C, kh, kw) if per else (O, C, kh, kw), device=dev) / (C * kh * kw) ** 0.5
    _C.conv_flags = _lib.CONV_FORCE_SIMT
    y = _C.conv2d_forward(x, w, 1, p)
    dy = torch.randn_like(y)
    dw = _C.conv2d_wgrad(dy, x, (kh, kw), 1, p, per)
    torch.cuda.synchronize()
    _C.conv_flags = _lib.CONV_FORCE_TC
    def markers():
        words = ctypes.c_size_t(0)
        ptr = _lib.lib().msg_debug_buffer(ctypes.byref(words))
        return [hex(int(v)) for v in np.ctypeslib.as_array(ptr, shape=(8,))] if ptr else None
    try:
        if which == "f":
            y2 = _C.conv2d_forward(x, w, 1, p); torch.cuda.synchronize()
            print(name, "fwd err", ((y2 - y).abs().max() / y.abs().max()).item(), markers(), flush=True)
        else:
            dw2 = _C.conv2d_wgrad(dy, x, (kh, kw), 1, p, per); torch.cuda.synchronize()
            print(name, "wgrad err", ((dw2 - dw).abs().max() / dw.abs().max()).item(), markers(), flush=True)
    except RuntimeError as e:
        print(name, which, "ERROR", str(e)[:80], markers(), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2], sys.argv[3]); sys.exit(0)
    for name in (sys.argv[1:] or CASES):
        for which in ("f", "w"):
            env = dict(os.environ, MSG_B200_TC_DEBUG=os.environ.get("MSG_B200_TC_DEBUG", "1"))
            try:
                r = subprocess.run([sys.executable, __file__, "child", name, which], env=env, timeout=120, capture_output=True, text=True)
                out = [l for l in r.stdout.splitlines() if l.startswith(name)]
                print("\n".join(out) if out else ("?? " + r.stderr[-300:]), flush=True)
            except subprocess.TimeoutExpired:
                print(name, which, "TIMEOUT", flush=True)
