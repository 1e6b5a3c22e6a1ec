"""Data path of the train step (first half): the reference's TIFF sequence dataset
(dataset/tlfm_dataset.py:15-198, dataset/utils.py:4-23) and a loader that keeps the GPU fed.

`TFLMDatasetGAN` mirrors the reference class: same constructor arguments, same sample index (position folders, z
positions, sort key, windows of one trap), `__getitem__` returns the same tensor bit for bit (tests/test_dataset.py against
outputs of the imported reference).  The per-sample work is decode + normalise; it stays on host threads.

`DeviceLoader` replaces `DataLoader(shuffle=True, drop_last=True, num_workers=B, pin_memory=True)` followed by
`.to(device)` in the training loop (train_multi_stylegan.py:60-63, model_wrapper.py:254-257):
  * worker threads decode straight into slots of a ring of PINNED batch slabs — no per-sample tensors, no collate copy, no
    pickling through worker processes (cv2 decoding releases the GIL);
  * each finished slab goes to its device twin with ONE asynchronous copy on a dedicated copy stream; the consumer's stream
    waits on the slab's event, so the upload of batch i+1 .. i+depth-1 overlaps the train step of batch i;
  * a slab is reused only after the consumer's stream has passed the point where the next batch is requested (event);
  * with several ranks every rank walks its own stride of ONE permutation that is seeded identically everywhere.
A batch of the benchmark's size is 8 x 2 x 3 x 256 x 256 fp32 = 12.6 MB: 0.2 ms over PCIe 5 against a 70 ms step; what
matters is that neither the decode nor the copy ever sits on the step's critical path."""
import os
import queue
import threading
from typing import Callable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch


def normalize_0_1(tensor: torch.Tensor, max: Optional[float] = None, min: Optional[float] = None) -> torch.Tensor:
    """dataset/utils.py:4-23: channel-wise (x - min) / (max - min) of a [channels, height, width] tensor."""
    flat = tensor.flatten(start_dim=1)
    lo = flat.min(dim=1, keepdim=True)[0].float() if min is None else torch.tensor(min, dtype=torch.float)
    hi = flat.max(dim=1, keepdim=True)[0].float() if max is None else torch.tensor(max, dtype=torch.float)
    return ((flat - lo) / (hi - lo)).reshape(tensor.shape)


def _read_image(path: str) -> np.ndarray:
    """cv2.imread(path, -1) of the reference (:143): the stored sample type (16-bit microscope TIFFs) unchanged."""
    try:
        import cv2
    except ImportError:                                  # same pixels through PIL where OpenCV is not installed
        from PIL import Image
        with Image.open(path) as im:
            return np.asarray(im)
    image = cv2.imread(path, -1)
    if image is None:
        raise FileNotFoundError("cannot decode %s" % path)
    return image


def _horizontal_flip_half(images: torch.Tensor) -> torch.Tensor:
    """The reference's default transformation, transforms.RandomHorizontalFlip(p=0.5) (:28-29): one torch.rand(1) draw."""
    return images.flip(-1) if torch.rand(1) < 0.5 else images


class TFLMDatasetGAN(torch.utils.data.Dataset):
    """Unsupervised trapped-yeast TLFM sequences for the generation task (dataset/tlfm_dataset.py:15-198).
    Returns [channels (BF, GFP, RFP as enabled), sequence_length, height, width] float32 in [0, 1]."""

    def __init__(self, path: str, sequence_length: int = 3, overlap: bool = True,
                 transformations: Optional[Callable[[torch.Tensor], torch.Tensor]] = _horizontal_flip_half,
                 z_position_indications: Tuple[str, ...] = ("_000_", "_001_", "_002_"),
                 gfp_min: Union[float, int] = 150.0, gfp_max: Union[float, int] = 2200.0,
                 rfp_min: Union[float, int] = 20.0, rfp_max: Union[float, int] = 2000.0, flip: bool = True,
                 positions: Optional[Tuple[str, ...]] = None, no_rfp: bool = False, no_gfp: bool = False) -> None:
        self.transformations = transformations if transformations is not None else (lambda x: x)
        self.gfp_min, self.gfp_max, self.rfp_min, self.rfp_max = gfp_min, gfp_max, rfp_min, rfp_max
        self.flip, self.no_rfp, self.no_gfp = flip, no_rfp, no_gfp
        self.sequence_length = sequence_length
        self.paths_to_dataset_samples: List[Tuple[Tuple[str, ...], Tuple[str, ...], Tuple[str, ...]]] = []
        step = 1 if overlap else sequence_length
        for position_folder in os.listdir(path):                                        # :57 (directory order, like the reference)
            folder = os.path.join(path, position_folder)
            if (positions is not None and position_folder not in positions) or not os.path.isdir(folder):
                continue
            files = [os.path.join(folder, f) for f in os.listdir(folder) if "tif" in f]  # :62-63
            per_channel = [self._by_z_position([f for f in files if tag in f], z_position_indications)
                           for tag in ("-BF0_", "-GFP", "-RFP")]                          # :65-69
            bf, gfp, rfp = per_channel
            for z in range(len(z_position_indications)):                                 # :101-110
                for index in range(0, len(bf[z]) - sequence_length + 1, step):
                    window = slice(index, index + sequence_length)
                    if self._check_if_same_trap(bf[z][window]):
                        self.paths_to_dataset_samples.append((tuple(bf[z][window]), tuple(gfp[z][window]),
                                                              tuple(rfp[z][window])))

    @staticmethod
    def _sort_key(item: str) -> str:
        """:76-78: time step (last `_` field of the last `-` part) followed by the fifth-last `_` field of the path."""
        return item.split("-")[-1].split("_")[-1].replace(".tif", "") + item.split("_")[-5]

    @classmethod
    def _by_z_position(cls, files: Sequence[str], indications: Sequence[str]) -> List[List[str]]:
        return [sorted((f for f in files if z in f), key=cls._sort_key) for z in indications]

    @staticmethod
    def _check_if_same_trap(path_list: Sequence[str]) -> bool:
        """:112-119: all paths name the same `trapNNNN`."""
        traps = [p[p.find("trap"):p.find("trap") + 8] for p in path_list]
        return all(t == traps[0] for t in traps)

    def __len__(self) -> int:
        return len(self.paths_to_dataset_samples)

    @property
    def channels(self) -> int:
        return 1 if self.no_gfp else (2 if self.no_rfp else 3)

    def _load(self, paths: Sequence[str]) -> torch.Tensor:
        return torch.stack([torch.from_numpy(_read_image(p).astype(np.float32)) for p in paths], dim=0)

    def __getitem__(self, item: int) -> torch.Tensor:
        """:128-198.  The channel selection follows the reference's precedence: no_gfp -> bright field only (whatever
        no_rfp says), else no_rfp -> BF + GFP, else all three."""
        paths_bf, paths_gfp, paths_rfp = self.paths_to_dataset_samples[item]
        n = self.channels
        if self.no_gfp and not self.no_rfp:          # the reference normalises images[2] of a 1-channel tensor here (:193-194)
            raise IndexError("index 2 is out of bounds for dimension 0 with size 1")
        stacks = [self._load(paths_bf)]
        if n >= 2:
            stacks.append(self._load(paths_gfp))
        if n == 3:
            stacks.append(self._load(paths_rfp))
        images = self.transformations(torch.cat(stacks, dim=0))                          # all frames as channels of one image
        images = images[0] if images.ndimension() == 4 else images
        images = images.reshape(n, images.shape[0] // n, *images.shape[1:]).clone()
        images[0] = normalize_0_1(images[0])                                             # :186, per frame
        if n >= 2:
            images[1] = ((images[1] - self.gfp_min).clamp(min=0.0) / self.gfp_max).clamp(max=1.0)   # :190
        if n == 3:
            images[2] = ((images[2] - self.rfp_min).clamp(min=0.0) / self.rfp_max).clamp(max=1.0)   # :194
        return images.flip(dims=(-2,)) if self.flip else images


class DeviceLoader(object):
    """Iterates a map-style dataset as device-resident batches [B, ...]: shuffled, last partial batch dropped, sharded over
    ranks; decode on `workers` host threads into pinned slabs, uploads on a copy stream `depth` batches ahead.

    The yielded tensor is a slot of a device ring: it stays valid until `depth - 1` further batches have been requested
    (ModelWrapper.train_step copies what it keeps).  `len()` = batches per epoch on this rank.  `set_epoch(e)` reseeds the
    permutation (every rank must use the same seed and epoch)."""

    def __init__(self, dataset, batch_size: int, device: Union[str, torch.device] = "cuda", shuffle: bool = True,
                 workers: int = 8, depth: int = 3, seed: int = 0, rank: int = 0, world_size: int = 1) -> None:
        if batch_size < 1 or depth < 2:
            raise ValueError("DeviceLoader: batch_size >= 1 and depth >= 2 (one slab uploading while one is consumed)")
        self.dataset, self.batch_size, self.device = dataset, int(batch_size), torch.device(device)
        self.shuffle, self.workers, self.depth = bool(shuffle), max(1, int(workers)), int(depth)
        self.seed, self.epoch, self.rank, self.world_size = int(seed), 0, int(rank), int(world_size)
        self._host: List[torch.Tensor] = []
        self._dev: List[torch.Tensor] = []

    def __len__(self) -> int:
        return (len(self.dataset) // self.world_size) // self.batch_size

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def batch_indices(self) -> List[List[int]]:
        """This rank's batches of the epoch: stride `world_size` through one permutation shared by all ranks."""
        n = len(self.dataset)
        if self.shuffle:
            order = torch.randperm(n, generator=torch.Generator().manual_seed(self.seed + self.epoch)).tolist()
        else:
            order = list(range(n))
        mine = order[self.rank:(n // self.world_size) * self.world_size:self.world_size]
        return [mine[i * self.batch_size:(i + 1) * self.batch_size] for i in range(len(mine) // self.batch_size)]

    def _slabs(self, sample: torch.Tensor) -> None:
        shape = (self.batch_size,) + tuple(sample.shape)
        if self._host and tuple(self._host[0].shape) == shape:
            return
        cuda = self.device.type == "cuda"
        self._host = [torch.empty(shape, dtype=sample.dtype, pin_memory=cuda) for _ in range(self.depth)]
        self._dev = [torch.empty(shape, dtype=sample.dtype, device=self.device) for _ in range(self.depth)]

    def __iter__(self) -> Iterator[torch.Tensor]:
        batches = self.batch_indices()
        if not batches:
            return
        first = self.dataset[batches[0][0]]
        self._slabs(first)
        cuda = self.device.type == "cuda"
        copy_stream = torch.cuda.Stream(self.device) if cuda else None
        ready = [torch.cuda.Event() for _ in range(self.depth)] if cuda else None       # upload of slot s finished
        released = [None] * self.depth               # consumer work on slot s's device twin (recorded on its stream)
        errors: "queue.Queue[BaseException]" = queue.Queue()
        tasks: "queue.Queue[Optional[Tuple[int, int, int, Optional[torch.Tensor]]]]" = queue.Queue()
        remaining = [0] * self.depth
        filled = [threading.Event() for _ in range(self.depth)]
        lock = threading.Lock()

        def work() -> None:
            while True:
                task = tasks.get()
                if task is None:
                    return
                slot, row, index, sample = task
                try:
                    self._host[slot][row].copy_(self.dataset[index] if sample is None else sample)
                except BaseException as exc:             # surfaced in the consumer, never swallowed
                    errors.put(exc)
                with lock:
                    remaining[slot] -= 1
                    if remaining[slot] == 0:
                        filled[slot].set()

        threads = [threading.Thread(target=work, daemon=True) for _ in range(min(self.workers, self.batch_size * 2))]
        for t in threads:
            t.start()

        def schedule(b: int) -> None:
            slot = b % self.depth
            filled[slot].clear()
            with lock:
                remaining[slot] = len(batches[b])
            for row, index in enumerate(batches[b]):
                tasks.put((slot, row, index, first if (b == 0 and row == 0) else None))

        try:
            for b in range(min(self.depth - 1, len(batches))):
                schedule(b)
            for b in range(len(batches)):
                slot = b % self.depth
                filled[slot].wait()
                if not errors.empty():
                    raise errors.get()
                if cuda:
                    current = torch.cuda.current_stream(self.device)
                    with torch.cuda.stream(copy_stream):
                        if released[slot] is not None:
                            copy_stream.wait_event(released[slot])       # the step that read this device slot has finished
                        self._dev[slot].copy_(self._host[slot], non_blocking=True)
                        ready[slot].record(copy_stream)
                    current.wait_event(ready[slot])
                else:
                    self._dev[slot].copy_(self._host[slot])
                nxt = b + self.depth - 1
                if nxt < len(batches):
                    # batch `nxt` takes the slot of batch b - 1, whose consumer work is everything issued so far
                    reuse = nxt % self.depth
                    if cuda:
                        released[reuse] = torch.cuda.Event()
                        released[reuse].record(torch.cuda.current_stream(self.device))
                        if b >= 1:
                            ready[reuse].synchronize()                   # its pinned slab has been uploaded: refill it
                    schedule(nxt)
                yield self._dev[slot]
        finally:
            for _ in threads:
                tasks.put(None)
            for t in threads:
                t.join()
