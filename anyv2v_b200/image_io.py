"""Frames in, frames out: the image / video plumbing either side of the two loops.

Mirrors what the reference runners and pipeline do around the hot path (none of it is inside "denoising-steps/sec"):
  * source frames ``{video_dir}/{video_name}/%05d.png`` (i2vgen-xl/utils.py:67-77), mp4 fallback (utils.py:43-64);
  * I2VGen-XL's first-frame geometry: "center crop wide" + bilinear resize to the CLIP crop
    (pipelines/pipeline_i2vgen_xl.py:1473-1509), VaeImageProcessor.preprocess with ``do_resize=False``
    (PIL -> float [-1, 1], :181);
  * export of the edited frames as png / gif / mp4 (run_group_pnp_edit.py:169-183; diffusers.utils.export_to_*).
Everything here is host-side PIL / numpy; nothing touches the GPU.
"""
from __future__ import annotations

import logging
import os
from typing import List, Sequence, Tuple

import numpy as np
import torch

logger = logging.getLogger(__name__)


def _pil():
    from PIL import Image
    return Image


def load_image(path: str):
    """diffusers.utils.load_image for local files: RGB, EXIF orientation applied."""
    from PIL import ImageOps
    img = _pil().open(path)
    img = ImageOps.exif_transpose(img)
    return img.convert("RGB")


def load_video_frames(frames_path: str, n_frames: int, image_size: Sequence[int] = (512, 512)):
    """i2vgen-xl/utils.py:67-77 — the first ``n_frames`` of ``%05d.png``; a size mismatch is a ValueError."""
    paths = [os.path.join(frames_path, f"{i:05d}.png") for i in range(n_frames)]
    frames = [load_image(p) for p in paths]
    want = tuple(int(v) for v in image_size)
    for f in frames:
        if tuple(f.size) != want:
            raise ValueError(f"Frame size {f.size} does not match config.image_size {want}")
    return paths, frames


def convert_video_to_frames(video_path: str, img_size: Sequence[int] = (512, 512), save_frames: bool = True):
    """i2vgen-xl/utils.py:43-64 — decode an mp4 into LANCZOS-resized PIL frames (optionally saved next to the video as
    ``{stem}/%05d.png``).  Video decoding is outside the hot path; OpenCV does it when it is importable."""
    try:
        import cv2
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("decoding a video file needs OpenCV (cv2); provide the frames as %05d.png instead") from e
    Image = _pil()
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {video_path}")
    out_dir = os.path.join(os.path.dirname(video_path), os.path.splitext(os.path.basename(video_path))[0])
    if save_frames:
        os.makedirs(out_dir, exist_ok=True)
    size = tuple(int(v) for v in img_size)
    frames = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        img = Image.fromarray(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
        if img.size != size:
            img = img.resize(size, resample=Image.Resampling.LANCZOS)
        if save_frames:
            img.save(os.path.join(out_dir, f"{len(frames):05d}.png"))
        frames.append(img)
    cap.release()
    return frames


def center_crop_wide(image, resolution: Tuple[int, int]):
    """pipeline_i2vgen_xl.py:1487-1509: shrink with a BOX filter until the image just covers ``resolution`` (w, h) —
    using the reference's ``round(size // scale)`` arithmetic — then crop the centre."""
    Image = _pil()
    many = isinstance(image, (list, tuple))
    imgs = list(image) if many else [image]
    rw, rh = int(resolution[0]), int(resolution[1])
    scale = min(imgs[0].size[0] / rw, imgs[0].size[1] / rh)
    out = []
    for u in imgs:
        u = u.resize((round(u.width // scale), round(u.height // scale)), resample=Image.BOX)
        left, top = (u.width - rw) // 2, (u.height - rh) // 2
        out.append(u.crop((left, top, left + rw, top + rh)))
    return out if many else out[0]


def resize_bilinear(image, resolution: Tuple[int, int]):
    """pipeline_i2vgen_xl.py:1473-1484."""
    Image = _pil()
    if isinstance(image, (list, tuple)):
        return [u.resize(tuple(resolution), Image.BILINEAR) for u in image]
    return image.resize(tuple(resolution), Image.BILINEAR)


def pil_to_unit_tensor(images) -> torch.Tensor:
    """PIL image(s) -> float32 [n, 3, H, W] in [0, 1] (VaeImageProcessor.pil_to_numpy + numpy_to_pt)."""
    imgs = list(images) if isinstance(images, (list, tuple)) else [images]
    arr = np.stack([np.asarray(u.convert("RGB"), dtype=np.float32) / 255.0 for u in imgs])
    return torch.from_numpy(arr).permute(0, 3, 1, 2).contiguous()


def preprocess(images) -> torch.Tensor:
    """VaeImageProcessor(do_resize=False, do_normalize=True).preprocess: PIL -> float32 [n, 3, H, W] in [-1, 1]."""
    return pil_to_unit_tensor(images) * 2.0 - 1.0


def frames_to_pil(video: torch.Tensor) -> List:
    """[f, 3, H, W] in [-1, 1] (one clip of decode_latents' output) -> list of PIL images (postprocess 'pil')."""
    Image = _pil()
    arr = ((video.float() / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).cpu().numpy() * 255).round().astype("uint8")
    return [Image.fromarray(a) for a in arr]


def export_to_gif(frames: Sequence, path: str, fps: int = 10) -> str:
    """diffusers.utils.export_to_gif: every frame shown 1000 / fps ms, looping forever."""
    frames[0].save(path, save_all=True, append_images=list(frames[1:]), optimize=False, duration=1000 // fps, loop=0)
    return path


def export_to_video(frames: Sequence, path: str, fps: int = 8) -> str:
    """diffusers.utils.export_to_video (OpenCV mp4v writer).  Without OpenCV the mp4 is skipped with a warning — the
    png frames and the gif still carry the result."""
    try:
        import cv2
    except ImportError:  # pragma: no cover
        logger.warning("cv2 not importable: %s not written", path)
        return ""
    w, h = frames[0].size
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    for f in frames:
        vw.write(cv2.cvtColor(np.asarray(f.convert("RGB")), cv2.COLOR_RGB2BGR))
    vw.release()
    return path
