import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = torch.ones(2, dtype=torch.float64, device=dev)
for _ in range(5): dist.all_reduce(t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): dist.all_reduce(t)
e1.record(); torch.cuda.synchronize()
if dist.get_rank() == 0: print(f"allreduce 16B: {e0.elapsed_time(e1)*10:.1f} us each (device time)", flush=True)
# with a busy compute stream: a long kernel occupying all SMs between all-reduces
a = torch.randn(8192, 8192, device=dev)
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        b = a @ a
        dist.all_reduce(t)
        c = t * 2
    torch.cuda.synchronize()
    if dist.get_rank() == 0: print(f"10 x (matmul + allreduce + use): {(time.perf_counter()-t0)*100:.2f} ms each", flush=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): b = a @ a
torch.cuda.synchronize()
if dist.get_rank() == 0: print(f"10 x matmul only: {(time.perf_counter()-t0)*100:.2f} ms each", flush=True)
dist.destroy_process_group()
