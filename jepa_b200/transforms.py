"""GPU input pipeline, host side: the reference's VideoTransform (app/vjepa/transforms.py:40-117) split
into (a) the RANDOM DECISIONS - crop box and flip, drawn here on the host in exactly the reference's RNG call order, so
a seeded run picks the same boxes - and (b) the PIXEL MATH, which runs in one CUDA kernel (csrc/preprocess.cu:
uint8 -> bilinear random-resized crop -> flip -> normalise -> [B,3,T,S,S]) after the uint8 frames have crossed PCIe.

    tf = make_transforms(crop_size=224, ...)          # same signature as the reference factory
    item = tf(buffer_uint8_THWC)                      # in the DataLoader worker: no pixel is touched, returns a ClipTicket
    clips = preprocess_batch([item, ...], device)     # on the training process: one H2D copy of uint8 frames + one kernel

auto_augment / motion_shift / random erasing are PIL / per-frame CPU augmentations outside this path; they raise.
"""
import math
import random

import numpy as np
import torch

from . import _lib
from . import kernels as K

DEFAULT_NORMALIZE = ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))


def get_param_spatial_crop(scale, ratio, height, width, num_repeat=10, log_scale=True, switch_hw=False):
    """Crop box (i, j, h, w) of a random-resized crop; RNG draws in the order of
    src/datasets/utils/video/transforms.py:503-542 (random.uniform x2, np.random.uniform, random.randint x2 per try)."""
    for _ in range(num_repeat):
        area = height * width
        target_area = random.uniform(*scale) * area
        if log_scale:
            log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
            aspect_ratio = math.exp(random.uniform(*log_ratio))
        else:
            aspect_ratio = random.uniform(*ratio)
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if np.random.uniform() < 0.5 and switch_hw:
            w, h = h, w
        if 0 < w <= width and 0 < h <= height:
            i = random.randint(0, height - h)
            j = random.randint(0, width - w)
            return i, j, h, w
    in_ratio = float(width) / float(height)     # fall back to a central crop
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


class ClipTicket:
    """One clip on its way to the GPU: the untouched uint8 frames [T,H,W,3] plus the decisions the kernel will apply."""
    __slots__ = ("frames", "box", "flip")

    def __init__(self, frames, box, flip):
        self.frames, self.box, self.flip = frames, box, flip


class GpuVideoTransform(object):
    """Drop-in for the reference's VideoTransform on its non-auto-augment path: __call__(buffer) consumes the same RNG
    draws (crop box, then one np.random.uniform for the flip, app/vjepa/transforms.py:100-108) and returns a ClipTicket."""

    def __init__(self, random_horizontal_flip=True, random_resize_aspect_ratio=(3 / 4, 4 / 3), random_resize_scale=(0.3, 1.0),
                 reprob=0.0, auto_augment=False, motion_shift=False, crop_size=224, normalize=DEFAULT_NORMALIZE):
        if auto_augment or motion_shift or reprob > 0:
            raise NotImplementedError("auto_augment / motion_shift / random erasing are CPU (PIL, per-frame) augmentations "
                                      "outside the GPU input path; every shipped pre-training config has them off")
        self.random_horizontal_flip = random_horizontal_flip
        self.random_resize_aspect_ratio = tuple(random_resize_aspect_ratio)
        self.random_resize_scale = tuple(random_resize_scale)
        self.crop_size = crop_size
        self.mean, self.std = tuple(normalize[0]), tuple(normalize[1])

    def __call__(self, buffer):
        frames = torch.as_tensor(np.ascontiguousarray(buffer))
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError("GpuVideoTransform expects decoded uint8 frames [T, H, W, 3]")
        box = get_param_spatial_crop(self.random_resize_scale, self.random_resize_aspect_ratio, frames.shape[1], frames.shape[2])
        flip = bool(np.random.uniform() < 0.5) if self.random_horizontal_flip else False
        return ClipTicket(frames, box, flip)


def make_transforms(random_horizontal_flip=True, random_resize_aspect_ratio=(3 / 4, 4 / 3), random_resize_scale=(0.3, 1.0),
                    reprob=0.0, auto_augment=False, motion_shift=False, crop_size=224, normalize=DEFAULT_NORMALIZE):
    """Same signature as app/vjepa/transforms.py:15-38."""
    return GpuVideoTransform(random_horizontal_flip=random_horizontal_flip, random_resize_aspect_ratio=random_resize_aspect_ratio,
                             random_resize_scale=random_resize_scale, reprob=reprob, auto_augment=auto_augment,
                             motion_shift=motion_shift, crop_size=crop_size, normalize=normalize)


def pack_tickets(tickets, pin=True):
    """Host staging of a batch: ONE contiguous (pinned) uint8 buffer with every clip's frames back to back and the
    [B, 10] int32 parameter table the kernel reads (byte offset lo/hi, H, W, i, j, h, w, flip, pad)."""
    sizes = [int(t.frames.numel()) for t in tickets]
    offs = np.concatenate([[0], np.cumsum([(s + 15) // 16 * 16 for s in sizes])]).astype(np.int64)
    buf = torch.empty(int(offs[-1]), dtype=torch.uint8, pin_memory=pin and torch.cuda.is_available())
    tab = np.zeros((len(tickets), 10), dtype=np.int32)
    for b, t in enumerate(tickets):
        buf[offs[b]:offs[b] + sizes[b]] = t.frames.reshape(-1)
        T, H, W, _ = t.frames.shape
        i, j, h, w = t.box
        tab[b, 0:2] = np.array([offs[b]], dtype=np.int64).view(np.int32)      # little-endian long long
        tab[b, 2:9] = (H, W, i, j, h, w, int(t.flip))
    table = torch.from_numpy(tab)
    if pin and torch.cuda.is_available():
        table = table.pin_memory()
    return buf, table


def preprocess_batch(tickets, device, crop_size, mean=DEFAULT_NORMALIZE[0], std=DEFAULT_NORMALIZE[1], dtype=torch.float32,
                     out=None):
    """uint8 tickets -> normalised clips [B, 3, T, S, S] on `device` (fp32 like the reference's loader, or bf16).  Host
    -> device traffic is the uint8 frames (T*H*W*3 bytes per clip) instead of fp32 crops (12*T*S*S bytes)."""
    import ctypes
    T = int(tickets[0].frames.shape[0])
    if any(int(t.frames.shape[0]) != T for t in tickets):
        raise ValueError("all clips of a batch must have the same number of frames")
    buf, table = pack_tickets(tickets)
    dbuf = buf.to(device, non_blocking=True)
    dtab = table.to(device, non_blocking=True)
    B, S = len(tickets), int(crop_size)
    if out is None:
        out = torch.empty(B, 3, T, S, S, dtype=dtype, device=device)
    m3 = (ctypes.c_float * 3)(*[float(x) for x in mean])
    s3 = (ctypes.c_float * 3)(*[float(x) for x in std])
    _lib.call("vj_clip_preprocess", dbuf.data_ptr(), dtab.data_ptr(), out.data_ptr(), 1 if out.dtype == torch.float32 else 0,
              B, T, S, ctypes.cast(m3, ctypes.c_void_p), ctypes.cast(s3, ctypes.c_void_p), K._s())
    return out     # dbuf / dtab go back to the caching allocator, which only reuses them in stream order
