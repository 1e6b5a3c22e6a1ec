"""Golden vectors for the data path: outputs of the UNMODIFIED reference dataset (dataset/tlfm_dataset.py, imported from
/root/reference) on a small synthetic TIFF tree.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_dataset          # writes tests/golden/dataset.pt  (needs /root/reference and cv2)

`synthetic_tree(root)` is also what tests/test_dataset.py calls to rebuild the same files (16-bit TIFFs written by cv2 from a
seeded generator), so the fixture only stores the reference's sample index and tensors, keyed by relative paths."""
import os
import sys

import numpy as np
import torch

POSITIONS = ("pos01", "pos02")
ZS = ("000", "001")
TRAPS = ((3, range(0, 5)), (11, range(4, 8)))        # (trap number, its time steps): windows across traps must be rejected
SIZE = (12, 10)

CONFIGS = {
    "bf_gfp": dict(no_rfp=True),
    "bf_gfp_rfp_no_overlap": dict(overlap=False, flip=False),
    "bf_only": dict(no_gfp=True, no_rfp=True, sequence_length=2),
    "one_position_len4": dict(no_rfp=True, positions=("pos02",), sequence_length=4, gfp_min=100.0, gfp_max=900.0),
}


def file_name(trap: int, channel: str, z: str, time: int) -> str:
    # fields chosen so that the reference's sort key (time step, then the fifth-last `_` field) and its `trapNNNN` test work
    return "x_trap%04d-%s_%s_w_e_%03d.tif" % (trap, channel, z, time)


def synthetic_tree(root: str) -> None:
    import cv2
    rng = np.random.default_rng(20240607)
    for pos in POSITIONS:
        os.makedirs(os.path.join(root, pos), exist_ok=True)
        for trap, times in TRAPS:
            for z in ZS:
                for t in times:
                    for channel, hi in (("BF0", 4000), ("GFP", 3000), ("RFP", 2500)):
                        image = rng.integers(0, hi, size=SIZE, dtype=np.uint16)
                        assert cv2.imwrite(os.path.join(root, pos, file_name(trap, channel, z, t)), image)
    with open(os.path.join(root, "notes.txt"), "w") as f:          # a non-directory entry next to the position folders
        f.write("not a position folder\n")


def rel(paths, root):
    return tuple(os.path.relpath(p, root) for p in paths)


def main() -> None:
    import tempfile
    sys.path.insert(0, os.environ.get("MSG_REFERENCE_ROOT", "/root/reference"))
    from torchvision import transforms
    import dataset as ref_dataset                                   # the reference package, unmodified
    out = {}
    with tempfile.TemporaryDirectory(prefix="msgds") as root:
        synthetic_tree(root)
        for name, kw in CONFIGS.items():
            ds = ref_dataset.TFLMDatasetGAN(path=root, z_position_indications=("_000_", "_001_"),
                                            transformations=transforms.Compose([]), **kw)
            out[name] = {"kwargs": kw,
                         "samples": {rel(s[0], root): (rel(s[1], root), rel(s[2], root), ds[i].clone())
                                     for i, s in enumerate(ds.paths_to_dataset_samples)}}
        # the default transformation (RandomHorizontalFlip) under a fixed seed: one draw per sample, in index order
        ds = ref_dataset.TFLMDatasetGAN(path=root, z_position_indications=("_000_", "_001_"), no_rfp=True, positions=("pos01",))
        torch.manual_seed(77)
        out["default_transform"] = {"order": [rel(s[0], root) for s in ds.paths_to_dataset_samples],
                                    "tensors": [ds[i].clone() for i in range(len(ds))]}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataset.pt")
    torch.save(out, path)
    print("wrote", path, {k: len(v.get("samples", v.get("order"))) for k, v in out.items()})


if __name__ == "__main__":
    main()
