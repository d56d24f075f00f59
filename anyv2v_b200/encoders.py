"""CLIP text / vision encoders of the I2VGen-XL pipeline — the once-per-clip steps in front of the two loops.

In the reference these are unchanged ``transformers`` modules held by the diffusers pipeline
(pipelines/pipeline_i2vgen_xl.py:142-166: ``tokenizer``, ``text_encoder`` = CLIPTextModel, ``image_encoder`` =
CLIPVisionModelWithProjection, ``feature_extractor`` = CLIPImageProcessor) and called from ``encode_prompt`` (:219-394)
and ``_encode_image`` (:395-425).  They are outside "denoising-steps/sec" and stay ``transformers`` modules here
(SURVEY 8f rank 4); this file only gives them the small surface the loops need:

    enc.encode_prompt("a red car")      -> [1, 77, D]   fp16, on the pipeline's device
    enc.encode_image(PIL first frame, width) -> [1, 1, D]   (center-crop-wide to a square, bilinear to the CLIP crop, CLIP mean/std)

``from_pretrained(dir)`` loads a local diffusers-layout checkpoint directory (``tokenizer/``, ``text_encoder/``,
``image_encoder/``, ``feature_extractor/``).  There is no network on the build / bench boxes, so ``random_init`` builds the
same architectures with seeded random weights and a self-contained deterministic tokenizer — enough to run the REAL input
path (PIL frames + prompt strings -> embeddings -> loops -> frames) end to end; the numbers mean nothing without weights.
"""
from __future__ import annotations

import re
import zlib
from typing import Sequence

import torch

from . import image_io

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

# OpenCLIP ViT-H/14, the towers of ali-vilab/i2vgen-xl (text width 1024 = the UNet's cross_attention_dim)
VIT_H = dict(text=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=23, num_attention_heads=16,
                       projection_dim=1024, vocab_size=49408, max_position_embeddings=77, hidden_act="gelu"),
             vision=dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
                         image_size=224, patch_size=14, projection_dim=1024, hidden_act="gelu"))


def tiny_arch(dim: int):
    """same topology, toy widths (CPU tests, smoke runs)"""
    return dict(text=dict(hidden_size=dim, intermediate_size=2 * dim, num_hidden_layers=2, num_attention_heads=2,
                          projection_dim=dim, vocab_size=4096, max_position_embeddings=77, hidden_act="gelu"),
                vision=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                            image_size=224, patch_size=14, projection_dim=dim, hidden_act="gelu"))


class HashTokenizer:
    """Offline stand-in for CLIPTokenizer (whose vocab / merges files cannot be downloaded here): lower-cased word and
    punctuation pieces are mapped to ids by CRC32, wrapped in BOS / EOS and padded to ``model_max_length`` like
    ``tokenizer(prompt, padding="max_length", truncation=True)`` (pipeline :283-289).  Deterministic across processes."""

    def __init__(self, vocab_size: int, model_max_length: int = 77):
        self.vocab_size, self.model_max_length = int(vocab_size), int(model_max_length)
        self.bos, self.eos = self.vocab_size - 2, self.vocab_size - 1

    def __call__(self, prompt, **_kw):
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        rows = []
        for p in prompts:
            pieces = re.findall(r"[a-z0-9]+|[^\sa-z0-9]", (p or "").lower())
            ids = [zlib.crc32(w.encode("utf-8")) % (self.vocab_size - 2) for w in pieces][: self.model_max_length - 2]
            ids = [self.bos] + ids + [self.eos]
            rows.append(ids + [self.eos] * (self.model_max_length - len(ids)))
        return torch.tensor(rows, dtype=torch.long)


class ClipEncoders:
    def __init__(self, tokenizer, text_encoder, image_encoder, crop_size=(224, 224), mean: Sequence[float] = CLIP_MEAN,
                 std: Sequence[float] = CLIP_STD, device="cpu", dtype=torch.float16):
        self.tokenizer = tokenizer
        self.device, self.dtype = torch.device(device), dtype
        self.text_encoder = text_encoder.to(device=self.device, dtype=dtype).eval()
        self.image_encoder = image_encoder.to(device=self.device, dtype=dtype).eval()
        self.crop_size = tuple(int(v) for v in crop_size)  # (width, height)
        self.mean = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
        self.std = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)

    # -- construction ------------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, model_dir: str, device="cpu", dtype=torch.float16):
        import os

        from transformers import CLIPImageProcessor, CLIPTextModel, CLIPTokenizer, CLIPVisionModelWithProjection
        sub = lambda s: os.path.join(model_dir, s)
        tok = CLIPTokenizer.from_pretrained(sub("tokenizer"))
        fe = CLIPImageProcessor.from_pretrained(sub("feature_extractor"))
        crop = fe.crop_size
        return cls(tok, CLIPTextModel.from_pretrained(sub("text_encoder")), CLIPVisionModelWithProjection.from_pretrained(sub("image_encoder")),
                   crop_size=(crop["width"], crop["height"]), mean=fe.image_mean, std=fe.image_std, device=device, dtype=dtype)

    @classmethod
    def random_init(cls, cross_dim: int = 1024, seed: int = 8888, device="cpu", dtype=torch.float16, arch=None):
        from transformers import CLIPTextConfig, CLIPTextModel, CLIPVisionConfig, CLIPVisionModelWithProjection
        arch = arch or (VIT_H if cross_dim == 1024 else tiny_arch(cross_dim))
        assert arch["text"]["hidden_size"] == cross_dim and arch["vision"]["projection_dim"] == cross_dim
        state = torch.random.get_rng_state()
        try:
            torch.manual_seed(seed)
            text = CLIPTextModel(CLIPTextConfig(**arch["text"]))
            vision = CLIPVisionModelWithProjection(CLIPVisionConfig(**arch["vision"]))
        finally:
            torch.random.set_rng_state(state)
        size = arch["vision"]["image_size"]
        tok = HashTokenizer(arch["text"]["vocab_size"], arch["text"]["max_position_embeddings"])
        return cls(tok, text, vision, crop_size=(size, size), device=device, dtype=dtype)

    # -- the two calls the pipeline makes -----------------------------------------------------------------------------
    @torch.no_grad()
    def encode_prompt(self, prompt) -> torch.Tensor:
        """pipeline :283-323 (clip_skip None, no attention mask): last hidden state of the text tower, [b, 77, D]."""
        if isinstance(self.tokenizer, HashTokenizer):
            ids = self.tokenizer(prompt)
        else:
            ids = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True,
                                 return_tensors="pt").input_ids
        return self.text_encoder(ids.to(self.device))[0].to(self.dtype)

    @torch.no_grad()
    def encode_image(self, image, width: int) -> torch.Tensor:
        """pipeline :1050-1054 + `_encode_image` :395-412: square center crop of side ``width``, bilinear resize to the CLIP
        crop, CLIP mean / std (no further resize / crop / rescale), projected image embedding -> [1, 1, D]."""
        sq = image_io.center_crop_wide(image, (width, width))
        sq = image_io.resize_bilinear(sq, self.crop_size)
        px = (image_io.pil_to_unit_tensor(sq) - self.mean) / self.std
        emb = self.image_encoder(px.to(device=self.device, dtype=self.dtype)).image_embeds
        return emb.unsqueeze(1).to(self.dtype)
