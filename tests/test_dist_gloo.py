"""CPU, world_size 2, gloo: the host-side logic of the image-sharded step -- shard ranges and the single
scalar exchange of the forward ELBO -- checked against the float64 oracle on the full batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_the_batch():
    from deepcgp_b200.dist import shard_range
    for n, w in [(256, 1), (256, 8), (512, 8), (10, 4), (7, 8)]:
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepcgp_b200.dist import elbo_from_partials, shard_range
    from oracle import dcgp_oracle as O
    from tests.util import layers_from_golden, load_golden
    g = load_golden("dgp2_elbo")
    layers = layers_from_golden(g)
    S, N = int(g["S"]), g["X"].shape[0]
    lo, hi = shard_range(N, rank, world)
    zs = [g["z%d" % i][:, lo:hi] for i in range(len(layers))]
    _, Fm, Fv = O.propagate(layers, g["X"][lo:hi], S, zs)
    K = Fm[-1].shape[2]
    ve = O.robustmax_varexp(Fm[-1].reshape(-1, K), Fv[-1].reshape(-1, K), np.tile(g["Y"][lo:hi].reshape(-1), S), K)
    part = torch.tensor([ve.sum()], dtype=torch.float64)
    kls = [O.layer_KL(l) for l in layers]
    elbo = elbo_from_partials(part, S, float(g["num_data"]), N, kls)
    q.put((rank, float(elbo.item())))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_elbo_equals_full_batch_oracle():
    from tests.util import load_golden
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    ref = float(load_golden("dgp2_elbo")["elbo"])
    for r in range(world):
        np.testing.assert_allclose(out[r], ref, rtol=1e-10)


def _bucket_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepcgp_b200.grad import Adam
    opt = Adam.__new__(Adam)                       # host logic only: no model, no device
    rng = np.random.RandomState(100 + rank)
    opt.grad = torch.tensor(rng.standard_normal(50))
    extra = torch.tensor([float(10 + rank)], dtype=torch.float64)
    before = opt.grad.clone()
    opt._allreduce_slice(5, 30, extra)             # one float32 bucket carrying the slice and the ELBO data term
    q.put((rank, before.numpy(), opt.grad.numpy().copy(), float(extra.item())))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_float32_gradient_bucket_with_elbo_term():
    """the gradient travels as ONE float32 bucket per slice; the ELBO's data term rides in its last slot."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {r[0]: r[1:] for r in (q.get(timeout=100) for _ in range(world))}
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    total = out[0][0] + out[1][0]
    for r in range(world):
        before, after, extra = out[r]
        np.testing.assert_allclose(after[5:30], total[5:30], rtol=1e-6, atol=1e-6)      # float32 exchange
        np.testing.assert_array_equal(after[:5], before[:5])                             # outside the slice: untouched
        np.testing.assert_array_equal(after[30:], before[30:])
        assert extra == 21.0
