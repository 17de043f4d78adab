"""Times adc_sample_assemble (device-side UniformSampling + input assembly) alone at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aerial_depth_completion_b200.dataloaders.dense_to_sparse import sample_assemble_cuda
B, H, W = int(os.environ.get("B", 32)), 228, 304
x = torch.rand(B, 4, H, W, device="cuda")
x[:, 3] = x[:, 3] * 9 + 1
seeds = torch.arange(B, dtype=torch.int32, device="cuda")
out = torch.empty(B, 5, H, W, device="cuda")
for _ in range(3):
    sample_assemble_cuda(x, seeds, 500, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    sample_assemble_cuda(x, seeds, 500, out=out)
e1.record()
torch.cuda.synchronize()
print("adc_sample_assemble B=%d %dx%d: %.1f us per call" % (B, H, W, e0.elapsed_time(e1) * 100))
