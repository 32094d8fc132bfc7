"""cuBLAS matmul throughput on this box (TF32 / BF16, 8192^3): the tensor-pipe denominators for the kNN roofline."""
import json
import torch

dev = torch.device("cuda:0")
out = {}
for name, dtype, tf32 in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False), ("fp32", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=dtype)
    b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    reps = 20 if name != "fp32" else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[name + "_tflops"] = 2.0 * n ** 3 / ms / 1e9
print(json.dumps(out))
