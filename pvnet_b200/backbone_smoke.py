"""Tiny backbone check for __graft_entry__.smoke(): native Resnet18_8s forward on a small image
against the PyTorch graph of the same module (fp32, TF32 off), plus the fused argmax."""
import torch


def run(dev="cuda:0"):
    from .model_repository import Resnet18_8s
    torch.manual_seed(0)
    net = Resnet18_8s(18, 2).to(dev).eval()
    x = torch.randn(2, 3, 96, 128, device=dev)
    with torch.no_grad():
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        ref = torch.cat(net._forward_torch(x), 1)
        torch.backends.cudnn.allow_tf32 = old
        out, mask = net.forward_native(x, with_mask=True)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-2, f"native backbone deviates from the torch graph: rel err {err:.3e}"
    assert torch.equal(mask, torch.argmax(out[:, :2], 1)), "fused argmax differs from torch.argmax of the logits"
    print(f"backbone smoke ok: rel err vs fp32 torch graph {err:.2e}")
