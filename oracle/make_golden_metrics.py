"""Golden vectors for the validation metrics' statistics: outputs of the UNMODIFIED reference's `FID._calc_fid` /
`FVD._calc_fvd` (multi_stylegan/validation_metrics.py:192-219, 401-428) and of its `misc.normalize_*_batch`
(misc.py:216-235), imported from /root/reference.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_metrics          # writes tests/golden/metrics.pt

Two stand-ins make the import possible: an empty `kornia` module (only IS's preprocessing uses it, which is not called here)
and an adapter for `scipy.linalg.sqrtm(..., disp=False)` — the reference's scipy returned `(sqrtm, error_estimate)` for
`disp=False`, this image's scipy has dropped the argument and returns the matrix."""
import importlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MSG_REFERENCE_ROOT", "/root/reference")


def load_reference_metrics():
    pkg = types.ModuleType("multi_stylegan")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "multi_stylegan")]
    saved = {k: sys.modules.get(k) for k in ("multi_stylegan", "kornia")}
    sys.modules["multi_stylegan"] = pkg
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))
    try:
        vm = importlib.import_module("multi_stylegan.validation_metrics")
        misc = importlib.import_module("multi_stylegan.misc")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.startswith("multi_stylegan.")]:
            sys.modules.pop(k)
    import scipy.linalg
    vm.sqrtm = lambda a, disp=False: (scipy.linalg.sqrtm(a), None)
    return vm, misc


def cases():
    """Seeded activation sets [samples, features]: correlated features, shifted means, more features than samples
    (rank-deficient covariances, where sqrtm turns complex and the reference keeps the real part)."""
    rng = np.random.default_rng(7)
    out = {}
    mix = rng.normal(size=(24, 24))
    out["full_rank"] = (rng.normal(size=(200, 24)) @ mix, rng.normal(size=(160, 24)) @ mix * 1.2 + 0.4)
    out["identical"] = (out["full_rank"][0], out["full_rank"][0].copy())
    out["rank_deficient"] = (rng.normal(size=(12, 32)), rng.normal(size=(10, 32)) + 0.1)
    out["float32_activations"] = (rng.normal(size=(96, 16)).astype(np.float32), (rng.normal(size=(96, 16)) * 0.5 + 1).astype(np.float32))
    return out


def main() -> None:
    vm, misc = load_reference_metrics()
    golden = {"frechet": {}, "normalize": {}}
    for name, (real, fake) in cases().items():
        fid, fvd = float(vm.FID._calc_fid(real, fake)), float(vm.FVD._calc_fvd(real, fake))
        assert fid == fvd
        golden["frechet"][name] = fid
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(3, 3, 2, 8, 8, generator=gen) * 5 + 1
    golden["normalize"] = {"input": x, "zero_one": misc.normalize_0_1_batch(x.clone()), "minus_one_one": misc.normalize_m1_1_batch(x.clone())}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "metrics.pt")
    torch.save(golden, path)
    print("wrote", path, golden["frechet"])


if __name__ == "__main__":
    main()
