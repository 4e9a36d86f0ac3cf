"""Multi-process (world_size 2, gloo, CPU) test of the data-parallel host logic: row sharding and the
single all-reduce of (sum log_q, count).  Per-rank log_q comes from the oracle (allowed in tests)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "normalizing-flows_b200"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_golden as lg
    from normflows.parallel import all_reduce_kld, shard_rows
    from oracle import nf_oracle as O
    spec, sd, a = lg("nsf_coupled_d5_h128_l3")
    x = a["x"].astype(np.float64)
    lo, hi = shard_rows(x.shape[0], rank, world)
    lp = O.log_prob(spec, sd, x[lo:hi])
    kld = all_reduce_kld(torch.tensor(lp.sum()), hi - lo)
    q.put((rank, float(kld), lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_forward_kld_all_reduce_matches_single_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, _, a = load_golden("nsf_coupled_d5_h128_l3")
    for _, kld, _, _ in res:
        assert kld == pytest.approx(float(a["kld_f64"]), rel=1e-6)  # golden kld accumulates in fp32
    spans = sorted((lo, hi) for _, _, lo, hi in res)
    assert spans[0][0] == 0 and spans[-1][1] == a["x"].shape[0]


def _grad_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "normalizing-flows_b200"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from normflows.parallel import GradientBuckets, shard_rows
    torch.manual_seed(0)  # same "model" on every rank
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((7, 5), (11,), (3, 4, 2), (1,))]
    frozen = torch.nn.Parameter(torch.randn(4), requires_grad=False)
    x = torch.randn(13, 5, generator=torch.Generator().manual_seed(1))  # 13 rows: ragged shards
    lo, hi = shard_rows(13, rank, world)

    def loss_of(rows):  # a mean over rows, like forward_kld
        h = rows @ params[0].t()
        return (h.pow(2).sum(1) + (h[:, :3] * params[2].sum((1, 2))).sum(1) + params[1].sum() * rows[:, 0]
                + params[3] * rows[:, 1]).mean()
    loss_of(x[lo:hi]).backward()
    gb = GradientBuckets(params + [frozen], bucket_bytes=200)  # tiny buckets: several collectives
    assert len(gb.buckets) >= 2
    gb.start(hi - lo).finish()
    got = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    loss_of(x).backward()  # single-process truth over the full batch
    err = max(float((g - p.grad).abs().max()) for g, p in zip(got, params))
    q.put((rank, err, frozen.grad is None))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gradient_buckets_average_like_one_process(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, err, frozen_untouched in res:
        assert err < 1e-5 and frozen_untouched


def _actnorm_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "normalizing-flows_b200"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import normflows as nf
    from normflows.parallel import shard_rows
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 6, generator=g) * torch.arange(1, 7) + 2.0           # flat batch, ragged shards
    img = torch.randn(10, 4, 3, 3, generator=g) * 0.5 - 1.0                  # image batch (batch_dims 0,2,3)
    out = {}
    for name, data, shape in (("flat", x, 6), ("img", img, (4, 1, 1))):
        lo, hi = shard_rows(data.shape[0], rank, world)
        for direction in ("inverse", "forward"):
            an = nf.flows.ActNorm(shape)
            an._data_init(data[lo:hi], direction)   # statistics only (the transform itself needs the GPU)
            out[f"{name}_{direction}"] = (an.s.detach().clone().numpy(), an.t.detach().clone().numpy())
    q.put((rank, out))
    dist.destroy_process_group()


def test_actnorm_init_is_global_batch_under_data_parallel():
    """with replicated parameters every rank must initialise ActNorm's s, t from the GLOBAL
    batch (all-reduce of sum x, sum x^2, n), identical on all ranks and equal to the single-process statistics
    (flows/normalization.py:19-39) of the unsharded batch."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + 17
    procs = [ctx.Process(target=_actnorm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path[:0] = [os.path.join(ROOT, "normalizing-flows_b200")]
    import normflows as nf
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 6, generator=g) * torch.arange(1, 7) + 2.0
    img = torch.randn(10, 4, 3, 3, generator=g) * 0.5 - 1.0
    for name, data, shape in (("flat", x, 6), ("img", img, (4, 1, 1))):
        for direction in ("inverse", "forward"):
            an = nf.flows.ActNorm(shape)
            an._data_init(data, direction)  # single process, whole batch
            for r in range(world):
                s, t = res[r][f"{name}_{direction}"]
                np.testing.assert_allclose(s, an.s.detach().numpy(), rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(t, an.t.detach().numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_array_equal(res[0][f"{name}_{direction}"][0], res[1][f"{name}_{direction}"][0])
            np.testing.assert_array_equal(res[0][f"{name}_{direction}"][1], res[1][f"{name}_{direction}"][1])
