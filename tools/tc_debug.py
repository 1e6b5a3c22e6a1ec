"""Diagnose the tcgen05 kernels: progress markers + smem tile dumps (MSG_B200_TC_DEBUG=1)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_dbg(n=256):
    import numpy as np
    from multi_stylegan_b200 import _lib
    words = ctypes.c_size_t(0)
    p = _lib.lib().msg_debug_buffer(ctypes.byref(words))
    if not p:
        return None
    return np.ctypeslib.as_array(p, shape=(words.value,))


def decode_tile(vals, label, rows=8):
    """values encode c*10000 + y*100 + x"""
    import numpy as np
    v = np.asarray(vals).view(np.float32).astype(np.int64)
    print(label, "first words (c,y,x) per 16B chunk:")
    for chunk in range(rows * 8):
        w = v[chunk * 4:(chunk + 1) * 4]
        print("  off %5d B:" % (chunk * 16), [(int(a // 10000), int(a % 10000 // 100), int(a % 100)) for a in w])


def child():
    import numpy as np
    import torch
    from multi_stylegan_b200 import _C, _lib
    dev = torch.device("cuda:0")
    which = os.environ.get("DBG_CASE", "pix32")
    print("case", which, "variant", os.environ.get("MSG_B200_TC_VARIANT", "0"), "tc:", _C.tensor_core_path_available(), flush=True)
    B, C, O, H, W = 1, 32, 64, 32, 32
    if which == "pix16":
        H = W = 16
    c = torch.arange(C, dtype=torch.float32).view(1, C, 1, 1)
    y = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    x = (c * 10000 + y * 100 + xx).to(dev)
    w = (torch.arange(O, dtype=torch.float32).view(O, 1) * 100 + torch.arange(C, dtype=torch.float32).view(1, C)).view(O, C, 1, 1).to(dev)
    xr = torch.randn(B, C, H, W, device=dev)
    wr = torch.randn(O, C, 1, 1, device=dev) / C ** 0.5
    _C.conv_flags = _lib.CONV_FORCE_SIMT
    ref = _C.conv2d_forward(xr, wr, 1, 0)
    dy = torch.randn_like(ref)
    ref_dw = _C.conv2d_wgrad(dy, xr, (1, 1), 1, 0, False)
    torch.cuda.synchronize()
    _C.conv_flags = _lib.CONV_FORCE_TC
    try:
        if which.startswith("pix"):
            out = _C.conv2d_forward(x, w, 1, 0)
            torch.cuda.synchronize()
            d = read_dbg()
            print("markers:", [hex(int(v)) for v in d[:8]], flush=True)
            decode_tile(d[256:256 + 4096], "A tile", rows=6)
            decode_tile(d[256 + 4096:256 + 4096 + 64 * 32], "B tile (n*100+c -> shown as (0,n,c))", rows=2)
            out2 = _C.conv2d_forward(xr, wr, 1, 0)
            torch.cuda.synchronize()
            print("markers(random run):", [hex(int(v)) for v in read_dbg()[:8]])
            err = ((out2 - ref).abs().max() / ref.abs().max()).item()
            print("forward rel err vs simt:", err, flush=True)
            if err > 1e-2:
                # how wrong? correlation per output channel / pixel
                print("out2[0,0,0,:8]", out2[0, 0, 0, :8].tolist())
                print("ref [0,0,0,:8]", ref[0, 0, 0, :8].tolist())
        else:
            dw = _C.conv2d_wgrad(dy, xr, (1, 1), 1, 0, False)
            torch.cuda.synchronize()
            d = read_dbg()
            print("markers:", [hex(int(v)) for v in d[:8]], flush=True)
            err = ((dw - ref_dw).abs().max() / ref_dw.abs().max()).item()
            print("wgrad rel err vs simt:", err, flush=True)
            if err > 1e-2:
                print("dw[0,:8]", dw[0, :8, 0, 0].tolist())
                print("ref[0,:8]", ref_dw[0, :8, 0, 0].tolist())
    except RuntimeError as e:
        print("ERROR", str(e)[:400], flush=True)
        d = read_dbg()
        if d is not None:
            print("markers after error:", [hex(int(v)) for v in d[:8]], flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    for case, variants in (("pix32", ["0", "1"]), ("pix16", ["0"]), ("red", ["0"])):
        for v in variants:
            env = dict(os.environ, MSG_B200_TC_VARIANT=v, MSG_B200_TC_DEBUG="1", DBG_CASE=case)
            print("=== case", case, "variant", v, flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, "child"], env=env, timeout=120, capture_output=True, text=True)
                print(r.stdout[-6000:], r.stderr[-1500:], "exit", r.returncode, flush=True)
            except subprocess.TimeoutExpired:
                print("TIMEOUT", flush=True)
