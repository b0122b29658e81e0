"""asr_mlp_forward (one fused launch) against the cuBLAS form of the same folded network, 8192 rows of 880 features."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import asr_b200 as A
from oracle import mlp_ref as mr
m = A.DenseStack(mr.random_weights(1))
x = torch.randn(8192, 880, device="cuda")
for name, fn in (("asr_mlp_forward (mlp_forward_kernel)", m.predict), ("cuBLAS addmm x6 + softmax", lambda t: torch.softmax(m.logits_library(t), 1))):
    for _ in range(3):
        fn(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{name}: {ms:.3f} ms per 8192 rows = {2 * 1.6e6 * 8192 / ms / 1e9:.1f} TFLOP/s")
