"""One training step of the flagship stack at the BASELINE batch (for `ncu --metrics gpu__time_duration.sum` launch
lists and CUDA-event phase timing): forward_kld + native backward."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "normalizing-flows_b200")]
import torch  # noqa: E402
import bench  # noqa: E402

B = int(os.environ.get("NFB_PROBE_BATCH", 65536))
model = bench.build_model("ar").cuda()
x = (torch.randn(B, 64, generator=torch.Generator().manual_seed(1)) * 1.5).cuda()
steps = int(os.environ.get("NFB_PROBE_STEPS", 3))
for i in range(steps):
    model.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = model.forward_kld(x)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: forward {1e3 * (t1 - t0):.2f} ms, backward {1e3 * (t2 - t1):.2f} ms, loss {float(loss):.4f}", flush=True)

# host cost of re-packing the weights (every optimizer step triggers one): nfb_flow_repack alone
from normflows._native import invalidate_packed_weights  # noqa: E402
for i in range(3):
    invalidate_packed_weights()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model._stack().ensure(64, x.device)
    torch.cuda.synchronize()
    print(f"repack {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
for i in range(3):
    model.zero_grad(set_to_none=True)
    loss = model.forward_kld(x)
    loss.backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    opt.step()
    torch.cuda.synchronize()
    print(f"adam step {i}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
