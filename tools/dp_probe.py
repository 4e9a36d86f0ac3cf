import os, sys, time, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "normalizing-flows_b200")]
import bench
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = torch.ones(2, dtype=torch.float64, device=dev)
for _ in range(20): dist.all_reduce(t)
model = bench.build_model().to(dev)
x = (torch.randn(65536, 64) * 1.5).to(dev)
h = model._stack()
buf = torch.zeros(2, dtype=torch.float64, device=dev)
for _ in range(5):
    h.forward_kld(x, sum_out=buf); dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier()
N = 30
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(N + 1)]
for i in range(N):
    ev[i][0].record()
    h.forward_kld(x, sum_out=buf)
    ev[i][1].record()
    dist.all_reduce(buf)
    ev[i][2].record()
ev[N][0].record()
torch.cuda.synchronize()
comp = [ev[i][0].elapsed_time(ev[i][1]) for i in range(N)]
ar = [ev[i][1].elapsed_time(ev[i][2]) for i in range(N)]
tot = ev[0][0].elapsed_time(ev[N][0]) / N
print(f"rank {dist.get_rank()}: step {tot:.3f} ms | compute mean {sum(comp)/N:.3f} max {max(comp):.3f} | "
      f"allreduce mean {sum(ar)/N*1e3:.1f} us max {max(ar)*1e3:.1f} us | first5 ar {[round(a*1e3) for a in ar[:5]]}", flush=True)
# same, but synchronise the host every step (no deep launch queue)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for i in range(N):
    h.forward_kld(x, sum_out=buf); dist.all_reduce(buf); torch.cuda.synchronize()
print(f"rank {dist.get_rank()}: host-synced step {(time.perf_counter()-t0)/N*1e3:.3f} ms", flush=True)
dist.destroy_process_group()
