"""Multi-GPU plumbing: one process per GPU, clips sharded one-per-rank, ONE collective — the weight broadcast.

The reference is single-process / single-GPU and loops over clips sequentially (run_group_pnp_edit.py:74); clips are
independent (every norm is per-sample), so they shard with no data-path collective.  The only exchange
is the one-time ``ncclBroadcast`` of the UNet's flat fp16 weight buffer (1.42 B params = 2.84 GB) from rank 0 over
NVLink / NVSwitch — ``torch.distributed`` is the plumbing (backend nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise the default process group from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def pick_device(configured: str):
    """Under torchrun every rank takes its LOCAL_RANK GPU; otherwise honour the config's ``device:`` string."""
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        _, local, _ = init_from_env()
        return torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    return torch.device(configured)


def shard_clips(n_items: int, rank: int, world: int) -> list[int]:
    """Round-robin clip -> rank assignment (clip i runs on rank i % world)."""
    return [i for i in range(n_items) if i % world == rank]


def flatten_parameters(module: torch.nn.Module) -> torch.Tensor:
    """Re-home every parameter/buffer of ``module`` as a view into ONE contiguous buffer (returned), so that the
    weight exchange is a single collective instead of ~1500 small ones."""
    tensors = [p for p in module.parameters()] + [b for b in module.buffers()]
    if not tensors:
        return torch.empty(0)
    dtype, device = tensors[0].dtype, tensors[0].device
    assert all(t.dtype == dtype and t.device == device for t in tensors), "flatten needs a uniform dtype/device"
    # keep every tensor 16-byte aligned inside the flat buffer (TMA descriptors and vector loads rely on it)
    align = 16 // tensors[0].element_size()
    offsets, total = [], 0
    for t in tensors:
        offsets.append(total)
        total += (t.numel() + align - 1) // align * align
    flat = torch.empty(total, dtype=dtype, device=device)
    with torch.no_grad():
        for t, off in zip(tensors, offsets):
            view = flat[off:off + t.numel()].view(t.shape)
            view.copy_(t)
            t.data = view
    return flat


def broadcast_flat(flat: torch.Tensor, src: int = 0) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)


def build_unet_replicated(cls, config: dict, seed: int, device, dtype=torch.float16, broadcast: bool = True,
                          checkpoint_dir: str | None = None):
    """Rank 0 draws the seeded random-init weights (zero-initialised TemporalConvLayer.conv4 re-randomised, std 0.02,
    like the oracle) — or loads ``checkpoint_dir`` (a diffusers ``unet/`` folder) when one is given; all ranks then hold
    bit-identical weights after one broadcast of the flat buffer."""
    rank, world = rank_world()
    state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        if rank == 0 or not broadcast or world == 1:
            net = cls(**config)
            g = torch.Generator().manual_seed(seed + 1)
            for name, m in net.named_modules():
                if name.endswith("conv4.3"):
                    with torch.no_grad():
                        m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.02)
                        m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.02)
            if checkpoint_dir is not None:
                from .run_group_pnp_edit import load_checkpoint_into
                if not load_checkpoint_into(net, checkpoint_dir):
                    raise FileNotFoundError(f"no diffusion_pytorch_model weights under {checkpoint_dir}")
        else:
            with torch.device("meta"):
                net = cls(**config)
            net = net.to_empty(device="cpu")
    finally:
        torch.random.set_rng_state(state)
    net = net.to(device=device, dtype=dtype).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    flat = flatten_parameters(net)
    net._broadcast_stats = None
    if broadcast and dist.is_initialized() and world > 1:
        net._broadcast_stats = timed_broadcast(flat, 0)
    net._flat_weights = flat
    return net


def timed_broadcast(flat: torch.Tensor, src: int = 0) -> dict:
    """The one collective of the path, timed: a 1-element warm-up broadcast first (NCCL builds its communicator and NVLink
    connections lazily on the first call), ranks aligned by a barrier, then the 2.84 GB transfer between two CUDA events."""
    import time
    nbytes = flat.numel() * flat.element_size()
    broadcast_flat(flat[:1], src)
    if flat.is_cuda:
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        broadcast_flat(flat, src)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        dist.barrier()
        t0 = time.perf_counter()
        broadcast_flat(flat, src)
        ms = (time.perf_counter() - t0) * 1e3
    return {"bytes": nbytes, "ms": round(ms, 3), "gb_per_s": round(nbytes / (ms * 1e-3) / 1e9, 1) if ms > 0 else None,
            "world": dist.get_world_size(), "backend": dist.get_backend(),
            "what": "dist.broadcast of the flat fp16 UNet weight buffer from rank 0 (outside the timed steps)"}
