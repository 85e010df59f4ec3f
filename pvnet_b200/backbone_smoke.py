"""Tiny backbone check for __graft_entry__.smoke(): native Resnet18_8s forward on a small image
against the PyTorch graph of the same module evaluated ON THE CPU in true fp32 (so the only GPU
kernels smoke() launches are the product's), plus the fused argmax."""
import torch


def run(dev="cuda:0"):
    from .model_repository import Resnet18_8s
    torch.manual_seed(0)
    net = Resnet18_8s(18, 2).eval()
    x = torch.randn(2, 3, 96, 128)
    with torch.no_grad():
        ref = torch.cat(net._forward_torch(x), 1)            # checker: torch CPU, fp32
        net = net.to(dev)
        out, mask = net.forward_native(x.to(dev), with_mask=True)
        out8, mask8 = net.forward_native(x.to(dev), with_mask=True, mask_dtype=torch.uint8)
    torch.cuda.synchronize()
    out, mask, out8, mask8 = out.cpu(), mask.cpu(), out8.cpu(), mask8.cpu()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-3, f"native backbone deviates from the fp32 torch graph: rel err {err:.3e}"   # TF32 class: ~2e-3
    assert torch.equal(mask, torch.argmax(out[:, :2], 1)), "fused argmax differs from torch.argmax of the logits"
    assert torch.equal(out, out8) and torch.equal(mask8.long(), mask)
    flips = (mask != torch.argmax(ref[:, :2], 1)).float().mean().item()
    assert flips < 1e-2, f"argmax flips vs the fp32 graph: {flips:.2e}"
    print(f"backbone smoke ok: rel err vs fp32 torch graph {err:.2e}, argmax flips {flips:.1e}")
