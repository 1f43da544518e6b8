"""Forward-only engine entry vs the training forward (DiT-XL/2, guided-sampling batch of 128): time and workspace.
Run on a B200: python scripts/dev_infer_fwd.py > profiles/r02f_infer_forward.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
from vaw_b200.models.dit import DiT_models  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda", 0)
    B = int(os.environ.get("B", 128))
    torch.manual_seed(0)
    m = DiT_models["DiT-XL"](image_size=32, patch_size=2, in_channels=4, class_dropout_prob=0.1, num_classes=1000,
                          learn_sigma=False).to(dev)
    x = torch.randn(B, 4, 32, 32, device=dev)
    t = torch.rand(B, device=dev) * 1000
    y = torch.randint(0, 1000, (B,), device=dev)
    flops = B * 711.7e9 / 3  # forward = 1/3 of the training FLOPs per image

    m.train()
    base = torch.cuda.memory_allocated()
    t_train = timed(lambda: m(x, t, y))
    ws_train = torch.cuda.memory_allocated() - base
    m.eval()
    with torch.no_grad():
        t_inf = timed(lambda: m(x, t, y))
    ws_inf = m._ws_inf.numel() * m._ws_inf.element_size() if m._ws_inf is not None else -1
    print(f"DiT-XL/2 forward, B = {B}")
    print(f"  training forward (stash for backward): {t_train:7.2f} ms  {flops / t_train / 1e9:6.0f} TFLOP/s  "
          f"workspace {ws_train / 2**30:.1f} GiB")
    print(f"  forward-only (torch.no_grad)         : {t_inf:7.2f} ms  {flops / t_inf / 1e9:6.0f} TFLOP/s  "
          f"workspace {ws_inf / 2**30:.2f} GiB")


if __name__ == "__main__":
    main()
