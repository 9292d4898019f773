"""Helpers shared by the -m gpu parity tests (inputs on device, step drivers)."""
import os
import zlib

import numpy as np
import torch

from oracle.inputs import logits_pool, rows_for  # noqa: F401


def load_case(golden_dir, cfg):
    data = np.load(os.path.join(golden_dir, cfg["name"] + ".npz"))
    pool = logits_pool(int(data["pool_seed"]), cfg["T"], cfg["V"], cfg["scale"])
    assert (zlib.crc32(pool.tobytes()) & 0xFFFFFFFF) == int(data["pool_crc"]), "input generator drifted"
    return data, pool


class PoolLogits:
    """logits_fn(t) -> [B, V] on device; stream r reads pool[(t + 3 r) % T] like oracle.inputs.rows_for."""

    def __init__(self, pool: np.ndarray, streams: int, device="cuda"):
        self.pool = torch.from_numpy(pool).to(device)
        self.T = pool.shape[0]
        self.offsets = 3 * torch.arange(streams, device=device)

    def __call__(self, t: int) -> torch.Tensor:
        idx = (t + self.offsets) % self.T
        return self.pool.index_select(0, idx).contiguous()
