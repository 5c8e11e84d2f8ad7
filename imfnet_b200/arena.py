"""Grow-only device arena: the fused plan's intermediates are bump-allocated from one buffer that is reused by every
forward, so a steady-state forward performs no cudaMalloc and no allocator search (sizes change with every fragment)."""
from __future__ import annotations

import torch


class Arena:
    def __init__(self, device, initial_bytes: int = 64 << 20):
        self.device = device
        self.buf = torch.empty(initial_bytes, dtype=torch.uint8, device=device)
        self.off = 0
        self._retired = []

    def reset(self):
        self.off = 0
        self._retired.clear()

    def take(self, nbytes: int) -> torch.Tensor:
        """uint8 view of `nbytes` (256-byte aligned)."""
        nbytes = max(int(nbytes), 1)
        start = (self.off + 255) // 256 * 256
        if start + nbytes > self.buf.numel():
            self._retired.append(self.buf)          # views handed out earlier in this forward stay valid
            self.buf = torch.empty(int((start + nbytes) * 1.5) + (16 << 20), dtype=torch.uint8, device=self.device)
            start = 0
        self.off = start + nbytes
        return self.buf[start:start + nbytes]

    def floats(self, rows: int, cols: int) -> torch.Tensor:
        rows = max(int(rows), 1)
        return self.take(rows * cols * 4).view(torch.float32).view(rows, cols)
