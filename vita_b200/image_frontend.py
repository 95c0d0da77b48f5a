"""Image front end on the GPU: decoded RGB bytes -> `[N, 3, 448, 448]` bf16 pixel values for InternViT.

Host-side mirror of the reference's two calls (video_audio_demo.py:214-221):
    image, p_num = dynamic_preprocess(image, min_num=1, max_num=12, image_size=448, use_thumbnail=True)
    image_tensor = model.process_images(image, model.config).to(dtype=model.dtype, device="cuda")
(`vita/util/data_utils_video_audio_neg_patch.py:1197-1255`, `vita/util/mm_utils.py:30-43`, CLIPImageProcessor with the
constants of `preprocessor_config.json`).  The tile-grid choice and Pillow's coefficient tables are computed here in
double precision exactly as Pillow does; the byte arithmetic (two resampling passes, tiling, rescale + normalise
through a 256-entry table) runs in csrc/image.cu and reproduces the reference's tensors bit for bit.  Decoding the
file (`Image.open(...).convert("RGB")`) stays on the host.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

from . import ops

IMAGE_SIZE = 448
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)
_PRECISION_BITS = 32 - 8 - 2


def closest_grid(width: int, height: int, min_num: int = 1, max_num: int = 12, image_size: int = IMAGE_SIZE) -> Tuple[int, int]:
    """Tile grid (columns, rows) with the aspect ratio closest to the image; ties go to the larger grid when the image
    has more than half its pixels (find_closest_aspect_ratio, data_utils...:1197-1211)."""
    grids = sorted({(i, j) for n in range(min_num, max_num + 1) for i in range(1, n + 1) for j in range(1, n + 1)
                    if min_num <= i * j <= max_num}, key=lambda g: g[0] * g[1])
    aspect = width / height
    best, best_diff = (1, 1), float("inf")
    for g in grids:
        diff = abs(aspect - g[0] / g[1])
        if diff < best_diff:
            best, best_diff = g, diff
        elif diff == best_diff and width * height > 0.5 * image_size * image_size * g[0] * g[1]:
            best = g
    return best


@lru_cache(maxsize=64)
def resample_tables(in_size: int, out_size: int):
    """Pillow's bicubic coefficients for resampling a full axis of `in_size` samples to `out_size`
    (Resample.c precompute_coeffs + normalize_coeffs_8bpc): (ksize, bounds [out, 2] int32, kk [out, ksize] int32).
    Vectorised over the output index; every floating-point operation keeps Pillow's order (the running sum of the
    weights is sequential over the taps), so the fixed-point values are identical."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5), 0.0)
    xmax = np.minimum(np.trunc(center + support + 0.5), float(in_size)) - xmin
    taps = np.arange(ksize, dtype=np.float64)[None, :]
    x = np.abs((taps + xmin[:, None] - center[:, None] + 0.5) * inv)
    a = -0.5
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    w = np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))
    w = np.where(taps < xmax[:, None], w, 0.0)
    ww = np.cumsum(w, axis=1)[:, -1:]                       # sequential, like `ww += w`
    k = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    fixed = np.trunc(np.where(k < 0, -0.5 + k * (1 << _PRECISION_BITS), 0.5 + k * (1 << _PRECISION_BITS)))
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return ksize, bounds, fixed.astype(np.int32)


def normalize_lut() -> torch.Tensor:
    """[3, 256] bf16: CLIPImageProcessor's rescale (uint8 -> float64 * (1/255) -> float32) and normalise
    ((x - mean) / std in float32), then the cast to the model dtype the demo applies."""
    v = (np.arange(256, dtype=np.float64) * (1 / 255)).astype(np.float32)
    mean = np.asarray(IMAGE_MEAN, dtype=np.float32)[:, None]
    std = np.asarray(IMAGE_STD, dtype=np.float32)[:, None]
    return torch.from_numpy(((v[None, :] - mean) / std).astype(np.float32)).to(torch.bfloat16)


def sample_frame_positions(n_frames: int, fps: float, max_frames: int, min_frames: int = 4, video_framerate: int = 1,
                           s=None, e=None):
    """Which decoded frames a video contributes (`_get_rawvideo_dec`, video_audio_demo.py:43-81): one frame per
    1 / video_framerate seconds inside [s, e], thinned to `max_frames` / repeated up to `min_frames` with
    `np.linspace(..., dtype=int)` exactly as the reference does.  Decoding itself (decord) stays on the host."""
    if s is None:
        start_time = end_time = None
    else:
        start_time, end_time = max(int(s), 0), max(int(e), 0)
        if start_time > end_time:
            start_time, end_time = end_time, start_time
        elif start_time == end_time:
            end_time = start_time + 1
    f_start = 0 if start_time is None else int(start_time * fps)
    f_end = int(min(1000000000 if end_time is None else end_time * fps, n_frames - 1))
    if f_end < f_start:
        return []
    stride = int(round(float(fps) / int(video_framerate)))
    all_pos = list(range(f_start, f_end + 1, stride))
    want = max_frames if len(all_pos) > max_frames else (min_frames if len(all_pos) < min_frames else None)
    if want is None:
        return all_pos
    return [all_pos[i] for i in np.linspace(0, len(all_pos) - 1, num=want, dtype=int)]


class _CudaBackend:
    """The two byte kernels of csrc/image.cu.  (Tests substitute a numpy stand-in to check the host-side geometry on
    machines without a GPU; the product never does.)"""

    def resample(self, img, axis, out_size, kk, bounds):
        return ops.image_resample_u8(img, axis, out_size, kk, bounds)

    def tiles_lut(self, img, lut, out, gi, gj, T, tile0):
        ops.image_tiles_lut(img, lut, out, gi, gj, T, tile0)


class ImageProcessor:
    """`preprocess(image)` = dynamic_preprocess + process_images of the reference in one device-side pipeline;
    `preprocess_frames(frames)` = the per-frame CLIP preprocessing of the video path."""

    def __init__(self, device="cuda", image_size: int = IMAGE_SIZE, backend=None):
        self.device = torch.device(device)
        self.image_size = image_size
        self.backend = backend or _CudaBackend()
        self.lut = normalize_lut().to(self.device)
        self.background = torch.tensor([int(x * 255) for x in IMAGE_MEAN], dtype=torch.uint8, device=self.device)
        self._tables = {}

    def _dev_tables(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        if key not in self._tables:
            _, bounds, kk = resample_tables(in_size, out_size)
            self._tables[key] = (torch.from_numpy(kk).to(self.device), torch.from_numpy(bounds).to(self.device))
            if len(self._tables) > 32:
                self._tables.pop(next(iter(self._tables)))
        return self._tables[key]

    def resize(self, img: torch.Tensor, width: int, height: int) -> torch.Tensor:
        """PIL `Image.resize((width, height))` (bicubic) on a device-resident [H, W, 3] uint8 image."""
        h, w, _ = img.shape
        out = img
        if width != w:
            kk, bounds = self._dev_tables(w, width)
            out = self.backend.resample(out, 1, width, kk, bounds)
        if height != h:
            kk, bounds = self._dev_tables(h, height)
            out = self.backend.resample(out, 0, height, kk, bounds)
        return out

    def to_device_u8(self, image) -> torch.Tensor:
        """PIL.Image / numpy [H, W, 3] uint8 / torch uint8 -> contiguous device tensor [H, W, 3]."""
        if hasattr(image, "convert") and hasattr(image, "size"):      # PIL image
            image = np.asarray(image.convert("RGB"))
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image))
        if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
            raise ValueError("expected an RGB image as [H, W, 3] uint8")
        return image.to(self.device).contiguous()

    def dynamic_tiles_u8(self, image, min_num: int = 1, max_num: int = 12, use_thumbnail: bool = True):
        """The resized full image and (if any) the thumbnail, still as bytes: (big [gj*T, gi*T, 3], thumb or None, (gi, gj))."""
        img = self.to_device_u8(image)
        h, w, _ = img.shape
        T = self.image_size
        gi, gj = closest_grid(w, h, min_num, max_num, T)
        big = self.resize(img, T * gi, T * gj)
        thumb = self.resize(img, T, T) if (use_thumbnail and gi * gj != 1) else None
        return big, thumb, (gi, gj)

    def preprocess(self, image, min_num: int = 1, max_num: int = 12, use_thumbnail: bool = True):
        """-> (pixel_values [N, 3, 448, 448] bf16 on the device, N) with the reference's tile order (+ thumbnail last)."""
        big, thumb, (gi, gj) = self.dynamic_tiles_u8(image, min_num, max_num, use_thumbnail)
        T = self.image_size
        n = gi * gj + (1 if thumb is not None else 0)
        out = torch.empty(n, 3, T, T, dtype=torch.bfloat16, device=self.device)
        self.backend.tiles_lut(big, self.lut, out, gi, gj, T, 0)
        if thumb is not None:
            self.backend.tiles_lut(thumb, self.lut, out, 1, 1, T, gi * gj)
        return out, n

    def process_tiles(self, tiles) -> torch.Tensor:
        """`process_images` for tiles that were already cut on the host (list of 448 x 448 PIL images / arrays)."""
        T = self.image_size
        out = torch.empty(len(tiles), 3, T, T, dtype=torch.bfloat16, device=self.device)
        for i, t in enumerate(tiles):
            u8 = self.to_device_u8(t)
            if u8.shape[0] != T or u8.shape[1] != T:      # CLIPImageProcessor would resize + centre-crop: not the tiled path
                raise ValueError("process_tiles expects 448 x 448 tiles (use preprocess() for whole images)")
            self.backend.tiles_lut(u8, self.lut, out, 1, 1, T, i)
        return out

    # -- video frames (video_audio_demo.py:83-110) -----------------------------------------------------------------
    def expand2square(self, img: torch.Tensor) -> torch.Tensor:
        """Pad the short side with int(mean * 255), image centred (mm_utils.py:16-28).  Byte copies only."""
        h, w, _ = img.shape
        if w == h:
            return img
        s = max(w, h)
        out = self.background.expand(s, s, 3).contiguous()
        if w > h:
            out[(w - h) // 2:(w - h) // 2 + h, :] = img
        else:
            out[:, (h - w) // 2:(h - w) // 2 + w] = img
        return out

    def clip_resize_crop(self, img: torch.Tensor) -> torch.Tensor:
        """CLIPImageProcessor geometry (transformers 4.41, PIL backend): shortest edge -> 448 with the long edge
        int(448 * long / short), bicubic, centre crop 448 x 448."""
        T = self.image_size
        h, w, _ = img.shape
        short, long = (w, h) if w <= h else (h, w)
        new_long = int(T * long / short)
        nh, nw = (new_long, T) if w <= h else (T, new_long)
        r = self.resize(img, nw, nh)
        top, left = (nh - T) // 2, (nw - T) // 2
        return r[top:top + T, left:left + T].contiguous()

    def preprocess_frames(self, frames, pad: bool = True) -> torch.Tensor:
        """Decoded video frames (sequence of [H, W, 3] uint8 / PIL images, or one [T, H, W, 3] array) ->
        [T, 3, 448, 448] bf16: expand2square (image_aspect_ratio == "pad") + CLIP resize / crop / rescale / normalise."""
        T = self.image_size
        frames = list(frames)
        out = torch.empty(len(frames), 3, T, T, dtype=torch.bfloat16, device=self.device)
        for i, f in enumerate(frames):
            u8 = self.to_device_u8(f)
            if pad:
                u8 = self.expand2square(u8)
            self.backend.tiles_lut(self.clip_resize_crop(u8), self.lut, out, 1, 1, T, i)
        return out
