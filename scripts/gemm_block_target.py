"""ncu target: one warm pass + one profiled pass over the twelve GEMM launches of a DiT-XL/2 block (bench.block_gemm_calls).
Usage under ncu:  ncu --set full -k regex:gemm_bf16_tcgen05 --launch-skip 12 -c 12 ... python scripts/gemm_block_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200")]
import torch
import bench
calls, flops = bench.block_gemm_calls(64, 1152, "cuda")
for _ in range(2):
    for c in calls:
        c()
    torch.cuda.synchronize()
print("done", sum(flops) / 1e12, "TFLOP per pass")
