"""GPU: nvJPEG front end (`pvnet_jpeg_decode_batch`) against PIL's libjpeg decode of the same bytes (the
reference's decoder, lib/datasets/linemod_dataset.py:180), and the decoded batch through the uint8 backbone
entry.  Different decoders: IDCT rounding (+-1..2 levels) for 4:4:4 files; for 4:2:0 files libjpeg's fancy chroma
upsampling differs at colour edges -- both measured and bounded here."""
import io

import numpy as np
import pytest
import torch

from pvnet_b200 import jpeg as pj

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pj.available(), reason="libnvjpeg not loadable")]
DEV = "cuda:0"


def _images(n, h=480, w=640):
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = []
    for i in range(n):
        img = np.stack([127 + 100 * np.sin(xx / (17 + i) + c) * np.cos(yy / (23 + 2 * i) - c) for c in range(3)], -1)
        img += rng.normal(0, 4, img.shape)
        out.append(np.clip(img, 0, 255).astype(np.uint8))
    return out


def _encode(img, subsampling):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=92, subsampling=subsampling)
    return buf.getvalue()


def _pil_decode(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("subsampling,max_bound,mean_bound", [(0, 6, 0.8), (2, 40, 1.5)], ids=["444", "420"])
def test_decode_matches_libjpeg_within_decoder_noise(subsampling, max_bound, mean_bound):
    imgs = _images(3)
    blobs = [_encode(im, subsampling) for im in imgs]
    dec = pj.JpegDecoder()
    got = dec.decode(blobs, 480, 640).cpu().numpy()
    ref = np.stack([_pil_decode(b) for b in blobs])
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    print(f"\n[nvJPEG vs libjpeg] subsampling {subsampling}: max |diff| {diff.max()}, mean {diff.mean():.3f}")
    assert got.shape == (3, 480, 640, 3)
    assert diff.max() <= max_bound and diff.mean() <= mean_bound
    again = dec.decode(blobs[:2], 480, 640)              # batch size change re-initialises
    assert torch.equal(again.cpu(), torch.from_numpy(got[:2]))


def test_decoded_batch_through_uint8_backbone_entry():
    from pvnet_b200.model_repository import Resnet18_8s
    from pvnet_b200.pipeline import IMAGENET_MEAN, IMAGENET_STD
    from tests.helpers import seeded_state_dict
    net = Resnet18_8s(18, 2)
    net.load_state_dict(seeded_state_dict(net, 3))
    net = net.to(DEV).eval()
    blobs = [_encode(im, 0) for im in _images(2, 96, 128)]
    img = pj.JpegDecoder().decode(blobs, 96, 128)
    with torch.no_grad():
        out = net.forward_native(img, mean=IMAGENET_MEAN, std=IMAGENET_STD)
        mean = torch.tensor(IMAGENET_MEAN, device=DEV).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, device=DEV).view(1, 3, 1, 1)
        ref = net.forward_native(img.permute(0, 3, 1, 2).float().div(255).sub(mean).div(std).contiguous())
    assert torch.equal(out, ref)


def test_wrong_size_and_garbage_are_errors():
    dec = pj.JpegDecoder()
    blob = _encode(_images(1, 96, 128)[0], 0)
    with pytest.raises(RuntimeError, match="expected"):
        dec.decode([blob], 480, 640)
    with pytest.raises(RuntimeError, match="JPEG"):
        dec.decode([b"not a jpeg at all"], 96, 128)
