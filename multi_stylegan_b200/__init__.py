"""multi_stylegan_b200 — B200-native (sm_100a) training-step hot path of Multi-StyleGAN.

Host code is Python/PyTorch and keeps the reference's API surface; all arithmetic on the hot path
runs in the hand-written CUDA library ``lib/libmsg_b200.so`` (C-ABI declared in ``include/msg_b200.h``).
There is no CPU fallback: every op raises if its input is not a CUDA tensor or the library is missing.
"""
__version__ = "0.1.0"

from ._mode import higher_order_gradients  # noqa: E402,F401
