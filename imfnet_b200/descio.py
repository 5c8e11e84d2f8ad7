"""Descriptor files: the `.npz{points, xyz, feature}` container scripts/generate_desc.py:118-123 writes and
scripts/evaluation_3dmatch.py:129-132 / scripts/benchmark_util.py:67-71 read.

The reference serialises the GPU here: `feature.detach().cpu().numpy()` (a blocking device->host copy) followed by zlib
compression on the calling thread, per fragment.  DescriptorWriter does the copy into pinned memory asynchronously on its own
stream and compresses on worker threads, so the next fragment's forward starts immediately; the files are byte-compatible
with numpy's loader (same keys, dtypes and shapes as the reference's)."""
from __future__ import annotations

import os
import queue
import threading

import numpy as np
import torch


def save_descriptors(path: str, points, xyz, feature) -> None:
    """Blocking equivalent of scripts/generate_desc.py:118-123."""
    if isinstance(feature, torch.Tensor):
        feature = feature.detach().cpu().numpy()
    np.savez_compressed(path, points=np.asarray(points), xyz=np.asarray(xyz), feature=np.asarray(feature))


def load_descriptors(path: str):
    """(points, xyz, feature) as scripts/benchmark_util.py:67-71 / evaluation_3dmatch.py:129-132 read them."""
    if not path.endswith(".npz"):
        path = path + ".npz"
    data = np.load(path)
    return data["points"], data["xyz"], data["feature"]


class DescriptorWriter:
    """Asynchronous writer: submit(path, points, xyz, feature_on_gpu) returns at once; close() waits for all files."""

    def __init__(self, workers: int = 2, device=None):
        self._q: "queue.Queue" = queue.Queue()
        self._err = []
        self._stream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(max(1, workers))]
        for t in self._threads:
            t.start()

    def submit(self, path: str, points, xyz, feature: torch.Tensor) -> None:
        if self._err:
            raise self._err[0]
        if isinstance(feature, torch.Tensor) and feature.is_cuda:
            host = torch.empty(feature.shape, dtype=feature.dtype, pin_memory=True)
            self._stream.wait_stream(torch.cuda.current_stream(feature.device))
            with torch.cuda.stream(self._stream):
                host.copy_(feature.detach(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._stream)
            feature.record_stream(self._stream)
            self._q.put((path, np.asarray(points), np.asarray(xyz), host, ev))
        else:
            self._q.put((path, np.asarray(points), np.asarray(xyz), torch.as_tensor(np.asarray(feature)), None))

    def _work(self):
        while True:
            item = self._q.get()
            if item is None:
                self._q.task_done()
                return
            path, points, xyz, host, ev = item
            try:
                if ev is not None:
                    ev.synchronize()
                os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
                np.savez_compressed(path, points=points, xyz=xyz, feature=host.numpy())
            except Exception as e:      # noqa: BLE001  (reported to the submitting thread)
                self._err.append(e)
            finally:
                self._q.task_done()

    def close(self) -> None:
        self._q.join()
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        if self._err:
            raise self._err[0]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
