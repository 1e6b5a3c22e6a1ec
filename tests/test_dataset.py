"""Data path: the TIFF sequence dataset against outputs of the unmodified reference dataset
(tests/golden/dataset.pt, made by oracle/make_golden_dataset.py on the same synthetic tree), and the device loader."""
import os

import pytest
import torch

from oracle import make_golden_dataset as mk
from tests.conftest import load_golden

Z = ("_000_", "_001_")


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("msgds"))
    assert "trap" not in root            # the reference finds the trap number by the first "trap" of the whole path
    mk.synthetic_tree(root)
    return root


def _rel(paths, root):
    return tuple(os.path.relpath(p, root) for p in paths)


def test_dataset_equals_the_reference_bit_for_bit(tree):
    from multi_stylegan_b200.dataset import TFLMDatasetGAN
    golden = load_golden("dataset.pt")
    for name, kw in mk.CONFIGS.items():
        ds = TFLMDatasetGAN(path=tree, z_position_indications=Z, transformations=None, **kw)
        want = golden[name]["samples"]
        assert len(ds) == len(want) > 0, name
        seen = set()
        for i, (bf, gfp, rfp) in enumerate(ds.paths_to_dataset_samples):
            key = _rel(bf, tree)
            w_gfp, w_rfp, w_tensor = want[key]
            assert _rel(gfp, tree) == w_gfp and _rel(rfp, tree) == w_rfp, (name, key)
            got = ds[i]
            assert got.dtype == torch.float32 and got.shape == w_tensor.shape, (name, got.shape, w_tensor.shape)
            assert torch.equal(got, w_tensor), (name, key)
            assert 0.0 <= float(got.min()) and float(got.max()) <= 1.0
            seen.add(key)
        assert seen == set(want)
        # within a position folder the reference's order (z position, then sorted windows) is kept
        by_pos = {}
        for bf, _, _ in ds.paths_to_dataset_samples:
            by_pos.setdefault(_rel(bf, tree)[0].split(os.sep)[0], []).append(_rel(bf, tree))
        ref_by_pos = {}
        for key in want:                                  # dict order = the reference's index order
            ref_by_pos.setdefault(key[0].split(os.sep)[0], []).append(key)
        assert by_pos == ref_by_pos, name


def test_default_transformation_consumes_the_random_stream_like_the_reference(tree):
    from multi_stylegan_b200.dataset import TFLMDatasetGAN
    golden = load_golden("dataset.pt")["default_transform"]
    ds = TFLMDatasetGAN(path=tree, z_position_indications=Z, no_rfp=True, positions=("pos01",))
    assert [_rel(s[0], tree) for s in ds.paths_to_dataset_samples] == golden["order"]
    torch.manual_seed(77)
    got = [ds[i] for i in range(len(ds))]
    flipped = 0
    for g, w in zip(got, golden["tensors"]):
        assert torch.equal(g, w)
    plain = TFLMDatasetGAN(path=tree, z_position_indications=Z, no_rfp=True, positions=("pos01",), transformations=None)
    flipped = sum(not torch.equal(plain[i], got[i]) for i in range(len(ds)))
    assert 0 < flipped < len(ds)                          # the seed draws both outcomes


def test_edge_cases(tree, tmp_path):
    from multi_stylegan_b200.dataset import TFLMDatasetGAN, normalize_0_1
    empty = TFLMDatasetGAN(path=str(tmp_path))
    assert len(empty) == 0
    too_long = TFLMDatasetGAN(path=tree, z_position_indications=Z, sequence_length=6, no_rfp=True)
    assert len(too_long) == 0                             # no trap has six time steps
    missing_z = TFLMDatasetGAN(path=tree, no_rfp=True)    # default indications name a z position the tree does not have
    assert len(missing_z) == 20
    with pytest.raises(IndexError):                       # reference: images[2] of a one-channel tensor (:193-194)
        TFLMDatasetGAN(path=tree, z_position_indications=Z, no_gfp=True, transformations=None)[0]
    x = torch.tensor([[[1.0, 3.0], [5.0, 9.0]], [[2.0, 2.5], [3.0, 4.0]]])
    n = normalize_0_1(x)
    assert float(n[0].min()) == 0.0 and float(n[0].max()) == 1.0 and float(n[1].max()) == 1.0
    assert torch.equal(normalize_0_1(x, max=9.0, min=1.0)[0], (x[0] - 1.0) / 8.0)


def _loader_checks(tree, device):
    from multi_stylegan_b200.dataset import DeviceLoader, TFLMDatasetGAN
    ds = TFLMDatasetGAN(path=tree, z_position_indications=Z, no_rfp=True, transformations=None)
    index = {tuple(ds[i].flatten()[:32].tolist()): i for i in range(len(ds))}
    epochs = []
    for world, rank in ((1, 0), (2, 0), (2, 1)):
        loader = DeviceLoader(ds, batch_size=3, device=device, workers=4, depth=3, seed=5, rank=rank, world_size=world)
        assert len(loader) == (len(ds) // world) // 3
        seen = []
        for batch in loader:
            assert batch.shape == (3, 2, 3) + mk.SIZE and batch.device.type == torch.device(device).type
            for row in batch.cpu():
                i = index[tuple(row.flatten()[:32].tolist())]
                assert torch.equal(row, ds[i])
                seen.append(i)
        assert len(seen) == len(set(seen)) == len(loader) * 3
        assert seen == [i for b in loader.batch_indices() for i in b]
        epochs.append(seen)
    assert not set(epochs[1]) & set(epochs[2])            # ranks read disjoint samples of one permutation
    loader = DeviceLoader(ds, batch_size=3, device=device, seed=5)
    first = [i for b in loader.batch_indices() for i in b]
    loader.set_epoch(1)
    assert first == epochs[0] and [i for b in loader.batch_indices() for i in b] != first
    ordered = DeviceLoader(ds, batch_size=4, device=device, shuffle=False)
    assert [i for b in ordered.batch_indices() for i in b] == list(range(20))

    class Broken(torch.utils.data.Dataset):
        def __len__(self):
            return 8

        def __getitem__(self, i):
            if i == 5:
                raise ValueError("sample 5 is unreadable")
            return torch.zeros(2, 2)
    with pytest.raises(ValueError):                       # a failing decode surfaces in the training loop
        for _ in DeviceLoader(Broken(), batch_size=2, device=device, shuffle=False):
            pass
    with pytest.raises(ValueError):
        DeviceLoader(ds, batch_size=2, device=device, depth=1)


def test_device_loader_host_logic(tree):
    _loader_checks(tree, "cpu")


@pytest.mark.gpu
def test_device_loader_uploads_on_the_copy_stream(tree):
    """On the device: same batches; and while the consumer keeps its stream busy the loader never hands out a slot whose
    previous contents are still being read (the consumer sums each batch late on a slow stream)."""
    from multi_stylegan_b200.dataset import DeviceLoader, TFLMDatasetGAN
    _loader_checks(tree, "cuda:0")
    ds = TFLMDatasetGAN(path=tree, z_position_indications=Z, no_rfp=True, transformations=None)
    loader = DeviceLoader(ds, batch_size=2, device="cuda:0", workers=4, depth=2, seed=1)
    want = [torch.stack([ds[i] for i in b]).double().sum() for b in loader.batch_indices()]
    busy = torch.randn(4096, 4096, device="cuda:0")
    got = []
    for batch in loader:
        for _ in range(3):
            busy = (busy @ busy).clamp(-1, 1)                # the "train step": keeps the consumer's stream behind the host
        got.append(batch.double().sum())                   # read AFTER the busy work, when the host is already batches ahead
    torch.cuda.synchronize()
    assert len(got) == len(want) == 10
    for g, w in zip(got, want):
        assert float(g.cpu()) == float(w)


class _Sequences(torch.utils.data.Dataset):
    """In-memory stand-in with the tiny networks' sample shape."""

    def __init__(self, n):
        gen = torch.Generator().manual_seed(4)
        self.data = torch.rand(n, 2, 3, 32, 32, generator=gen)
        self.reads = []

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        self.reads.append(i)
        return self.data[i]


def _train_over_loader(dev):
    """ModelWrapper.train (model_wrapper.py:104-145, 245-257) fed by the device loader: every epoch walks a new
    permutation, the last partial batch is dropped, one history entry per iteration."""
    from multi_stylegan_b200.dataset import DeviceLoader
    from multi_stylegan_b200.model_wrapper import ModelWrapper
    from tests.test_train_step import _hp, build
    G, D = build(dev)
    hp = _hp()
    opt_g = torch.optim.Adam(G.get_parameters(lr_main=2e-3, lr_style=2e-5), betas=hp["betas"])
    opt_d = torch.optim.Adam(D.parameters(), lr=6e-3, betas=hp["betas"])
    ds = _Sequences(7)
    loader = DeviceLoader(ds, batch_size=2, device=dev, workers=2, depth=2, seed=3)
    before = [p.detach().clone() for p in G.parameters()]
    mw = ModelWrapper(G, D, opt_g, opt_d, training_dataset=loader, hyperparameters=hp, device=dev)
    history = mw.train(epochs=2)
    assert len(history) == 2 * 3 and mw.iteration == 6
    assert all(torch.isfinite(v).all() for h in history for v in h.values())
    assert "loss_path_length_regularization" in history[1] and "loss_discriminator_regularization" in history[1]
    first, second = ds.reads[:len(ds.reads) // 2], ds.reads[len(ds.reads) // 2:]
    assert len(first) == len(second) == 6 and len(set(first)) == 6 and first != second
    assert any(not torch.equal(a, b) for a, b in zip(before, G.parameters()))


def test_training_loop_over_the_device_loader_host_logic(oracle_backend):
    _train_over_loader("cpu")


@pytest.mark.gpu
def test_training_loop_over_the_device_loader(built_library):
    _train_over_loader(torch.device("cuda:0"))
