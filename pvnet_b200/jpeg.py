"""JPEG bytes -> uint8 [b,H,W,3] CUDA tensor via nvJPEG (`pvnet_jpeg_decode_batch`), the front of
`Resnet18_8s.forward_native(uint8 images, mean=, std=)`.  Stands in for the reference's host-side
`Image.open(...)` + `ToTensor` + `Normalize` (lib/datasets/linemod_dataset.py:180-195, tools/demo.py:89-95):
compressed bytes are the only thing that crosses PCIe."""
from __future__ import annotations

import ctypes
import weakref

import torch

from . import _native


def available() -> bool:
    return bool(_native.lib().pvnet_jpeg_available())


class JpegDecoder:
    def __init__(self):
        self._h = ctypes.c_void_p()
        _native.check(_native.lib().pvnet_jpeg_decoder_create(ctypes.byref(self._h)), "pvnet_jpeg_decoder_create")
        self._fin = weakref.finalize(self, JpegDecoder._destroy, self._h.value)

    @staticmethod
    def _destroy(h):
        try:
            _native.lib().pvnet_jpeg_decoder_destroy(ctypes.c_void_p(h))
        except Exception:
            pass

    def decode(self, jpegs, height, width, device=None, out=None):
        """jpegs: sequence of `bytes` (one JPEG file each, all height x width) -> uint8 [b,height,width,3] RGB on
        `device` (current CUDA device by default), enqueued on the current stream."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        b = len(jpegs)
        if out is None:
            out = torch.empty([b, height, width, 3], dtype=torch.uint8, device=dev)
        ptrs = (ctypes.c_char_p * b)(*[bytes(j) for j in jpegs])
        lens = (ctypes.c_size_t * b)(*[len(j) for j in jpegs])
        with torch.cuda.device(dev):
            _native.check(_native.lib().pvnet_jpeg_decode_batch(
                self._h, ptrs, lens, b, int(height), int(width), out.data_ptr(),
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "pvnet_jpeg_decode_batch")
        return out
