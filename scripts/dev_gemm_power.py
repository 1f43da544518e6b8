"""Is the GEMM epilogue cost a time cost or an energy cost?  Runs the fc1-GELU GEMM back to back for ~2 s per setting and
samples SM clock / power (NVML) meanwhile.  VAW_DBG is read per process, so run once per setting."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "variance-aware-weight_b200"), os.path.join(ROOT, "tests")]
import torch, pynvml
from vaw_b200 import _lib as L
from gpu_util import run_gemm
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda"; M, D = 16384, 1152; N, K = 4 * D, D
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
A = bf(M, K); W = bf(N, K); o1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16); o2 = torch.empty_like(o1)
bias = torch.zeros(N, device=dev)
fn = lambda: run_gemm(A, W, 0, 0, M, N, K, L.EPI_GELU_TANH, out=o1, out2=o2, bias=bias, tile_n=256, cta_group=2)
for _ in range(20): fn()
torch.cuda.synchronize()
samples = []; stop = False
def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.02)
th = threading.Thread(target=poll); th.start()
iters = int(os.environ.get("ITERS", 8000))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters): fn()
e.record(); torch.cuda.synchronize()
stop = True; th.join()
us = s.elapsed_time(e) / iters * 1e3
tail = samples[len(samples) // 2:]
print(f"VAW_DBG={os.environ.get('VAW_DBG','0'):>6s}: {us:7.1f} us/launch  {2*M*N*K/us/1e6:7.1f} TF  sm_clock {sum(c for c,_ in tail)/len(tail):6.0f} MHz  power {sum(p for _,p in tail)/len(tail):6.0f} W  ({len(samples)} samples)")
