"""Small workload for compute-sanitizer (racecheck / synccheck / memcheck) on the persistent whole-stack kernel:
8 x [AR spline block + LU] and 8 x [coupled block + LU] (density AND sampling direction), 1024 + 37 rows, so that
(layer, tile) units of different layers really overlap across CTAs and the progress-flag protocol is exercised."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "normalizing-flows_b200")]
import torch  # noqa: E402
import bench  # noqa: E402

torch.set_grad_enabled(False)
x = (torch.randn(1024 + 37, 64, generator=torch.Generator().manual_seed(3)) * 1.5).cuda()
for kind in ("ar", "coupled"):
    m = bench.build_model(kind, layers=8).cuda()
    lp = m.log_prob(x)
    assert len(m._stack().fused_layers()) == 16
    print(kind, "log_prob ok", float(lp.mean()), "launches", m._stack().launch_count())
    if kind == "coupled":
        z, _ = m.inverse_and_log_det(x)
        xr, _ = m.forward_and_log_det(z)
        print("coupled round trip max err", float((xr - x).abs().max()))
torch.cuda.synchronize()
print("done")
