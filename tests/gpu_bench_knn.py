"""Timing driver (not a test): exhaustive Hamming search, config C3's unit (64 x 64 keyframes of 2000 descriptors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dvmslam_b200 import synth
from dvmslam_b200.matching import HammingKnn

KA = int(sys.argv[1]) if len(sys.argv) > 1 else 64
KB = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = 2000
A = torch.from_numpy(synth.keyframe_blocks(KA, N, seed=1)).cuda()
B = torch.from_numpy(synth.keyframe_blocks(KB, N, seed=2)).cuda()
k1 = torch.empty((KA, KB, N), dtype=torch.int32, device="cuda"); k2 = torch.empty_like(k1)
cnt = torch.empty((KA, KB), dtype=torch.int32, device="cuda")
h = HammingKnn(stream=torch.cuda.current_stream().cuda_stream)
MODE = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h.set_mode(MODE)
def run():
    h.knn_device(A.data_ptr(), KA, N, B.data_ptr(), KB, N, k1.data_ptr(), k2.data_ptr(), cnt.data_ptr(), 50, 0.75)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
pairs = KA * KB * N * N
print(f"mode {MODE}: {2 * pairs * 256 / ms * 1e-9:.1f} Tops/s (int8 MAC x2)")
print(f"knn {KA}x{KB} keyframes x {N}^2: {ms:.3f} ms  {pairs / ms / 1e6:.1f} G descriptor pairs/s  "
      f"{pairs * 8 / ms / 1e6 / 148 / 1.965:.2f} popc32/clk/SM  {KA * KB / ms * 1e3:.0f} keyframe pairs/s")
# single keyframe pair (split path)
k1s = torch.empty((1, 1, N), dtype=torch.int32, device="cuda"); k2s = torch.empty_like(k1s)
for _ in range(3): h.knn_device(A.data_ptr(), 1, N, B.data_ptr(), 1, N, k1s.data_ptr(), k2s.data_ptr())
torch.cuda.synchronize(); e0.record()
for _ in range(20): h.knn_device(A.data_ptr(), 1, N, B.data_ptr(), 1, N, k1s.data_ptr(), k2s.data_ptr())
e1.record(); torch.cuda.synchronize()
print(f"knn 1x1 keyframes: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
