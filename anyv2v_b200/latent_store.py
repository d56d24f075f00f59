"""Inverted-latent hand-off between the two phases.

The reference exchanges latents through the filesystem: ``invert`` does a D2H copy + ``torch.save`` of
``ddim_latents_{t}.pt`` EVERY step (pipeline_i2vgen_xl.py:1422-1428) and ``sample_with_pnp`` does a ``torch.load`` +
H2D EVERY step (:1134, i2vgen-xl/utils.py:25-30), each a host sync inside the hot loop.  Here all T latents stay
resident in HBM (50 x 0.5 MB for 16 f x 512^2), and the same on-disk files — same names, same ``torch.save`` payload
of a [1,4,F,h,w] fp16 tensor — are written by a background thread from pinned host copies, so phase 2 of a later
process (or the unmodified reference) can still pick them up.
"""
from __future__ import annotations

import glob
import os
import queue
import threading

import torch


def latent_path(ddim_latents_path: str, t) -> str:
    return os.path.join(ddim_latents_path, f"ddim_latents_{int(t)}.pt")


def load_ddim_latents_at_t(t, ddim_latents_path, map_location=None):
    """i2vgen-xl/utils.py:25-30, plus ``map_location`` (the reference's files are device-coupled)."""
    path = latent_path(ddim_latents_path, t)
    assert os.path.exists(path), f"Missing latents at t {t} path {path}"
    return torch.load(path, map_location=map_location)


def load_ddim_latents_at_T(ddim_latents_path, map_location=None):
    """i2vgen-xl/utils.py:33-39 — the noisiest saved latent."""
    ts = [int(os.path.basename(p).split("_")[-1].split(".")[0]) for p in glob.glob(os.path.join(ddim_latents_path, "ddim_latents_*.pt"))]
    assert ts, f"no ddim_latents_*.pt under {ddim_latents_path}"
    return load_ddim_latents_at_t(max(ts), ddim_latents_path, map_location)


class LatentStore:
    """{timestep -> latent} resident on the device, with optional asynchronous reference-format files."""

    def __init__(self, output_dir: str | None = None, write_files: bool = True, host_resident: bool = False):
        """host_resident=True keeps the latents in PINNED HOST memory instead of HBM (what the reference's disk
        hand-off amounts to): ``put`` is an async D2H copy, ``get`` an async H2D copy on the current stream."""
        self.output_dir = output_dir
        self.host_resident = host_resident
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.write_files = bool(write_files and output_dir)
        self._mem: dict[int, torch.Tensor] = {}
        self._q: queue.Queue | None = None
        self._worker: threading.Thread | None = None
        self._errors: list[BaseException] = []

    # -- device side ------------------------------------------------------------------------------------------------
    def put(self, t, latents: torch.Tensor) -> None:
        t = int(t)
        if self.host_resident and latents.is_cuda:
            keep = torch.empty(latents.shape, dtype=latents.dtype, pin_memory=True)
            keep.copy_(latents.detach(), non_blocking=True)  # stream-ordered D2H; consumed by a later stream-ordered H2D
            self.d2h_bytes += keep.numel() * keep.element_size()
        else:
            keep = latents.detach().clone()
        self._mem[t] = keep
        if self.write_files:
            self._enqueue(t, keep)

    def get(self, t, device=None) -> torch.Tensor:
        t = int(t)
        if t in self._mem:
            x = self._mem[t]
            if device is None or x.device == torch.device(device):
                return x
            if x.is_pinned():
                self.h2d_bytes += x.numel() * x.element_size()
            return x.to(device, non_blocking=True)
        if self.output_dir is None:
            raise KeyError(f"no inverted latent for t={t}")
        x = load_ddim_latents_at_t(t, self.output_dir, map_location=device or "cpu")
        self._mem[t] = x
        return x

    def __contains__(self, t) -> bool:
        return int(t) in self._mem or (self.output_dir is not None and os.path.exists(latent_path(self.output_dir, t)))

    def timesteps(self):
        return sorted(self._mem)

    # -- file writer ------------------------------------------------------------------------------------------------
    def _enqueue(self, t: int, x: torch.Tensor) -> None:
        if self._worker is None:
            os.makedirs(self.output_dir, exist_ok=True)
            self._q = queue.Queue()
            self._worker = threading.Thread(target=self._drain, name="latent-writer", daemon=True)
            self._worker.start()
        if not x.is_cuda and x.is_pinned():
            ev = torch.cuda.Event()
            ev.record()  # the D2H that fills this pinned buffer is ordered before this event
            host = x
        elif x.is_cuda:
            host = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            host.copy_(x, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host, ev = x, None
        self._q.put((t, host, ev))

    def _drain(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            t, host, ev = item
            try:
                if ev is not None:
                    ev.synchronize()
                tmp = latent_path(self.output_dir, t) + ".tmp"
                torch.save(host.clone(), tmp)
                os.replace(tmp, latent_path(self.output_dir, t))
            except BaseException as e:  # surfaced by flush()
                self._errors.append(e)

    def flush(self) -> None:
        """Block until every queued file is on disk (called once after the loop, never inside it)."""
        if self._worker is not None:
            self._q.put(None)
            self._worker.join()
            self._worker = None
        if self._errors:
            raise RuntimeError(f"latent writer failed: {self._errors[0]!r}")
