"""World-size-2 CPU (gloo) tests of the data-parallel host logic: batch sharding and the bucketed gradient
all-reduce must reproduce the single-process large-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from svit_b200.distributed import GradAllReducer, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(16, 32), nn.GELU(), nn.Linear(32, 32), nn.LayerNorm(32), nn.Linear(32, 5),
                         nn.Linear(5, 5))  # last layer unused below -> exercises the "no gradient" path


def _worker(rank, world, port, bucket_bytes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _net()
        red = GradAllReducer(net.parameters(), bucket_bytes=bucket_bytes)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(8, 16, generator=g)
        y = torch.randint(0, 5, (8,), generator=g)
        idx = shard_batch(8, rank, world)
        xs, ys = x[idx.start:idx.stop], y[idx.start:idx.stop]
        for _ in range(2):  # two steps: buckets must reset correctly
            for p in net.parameters():
                p.grad = None
            red.prepare()
            loss = nn.functional.cross_entropy(net[:5](xs), ys)
            loss.backward()
            red.finish()
        if rank == 0:
            torch.save([p.grad.clone() for p in net.parameters()], out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [1 << 30, 1024])
def test_bucketed_allreduce_equals_large_batch(tmp_path, bucket_bytes):
    out = str(tmp_path / "grads.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, bucket_bytes, out), nprocs=2, join=True)
    got = torch.load(out)
    net = _net()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 16, generator=g)
    y = torch.randint(0, 5, (8,), generator=g)
    nn.functional.cross_entropy(net[:5](x), y).backward()
    for p, gg in zip(net.parameters(), got):
        want = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(gg, want, rtol=1e-5, atol=1e-7)


def test_shard_batch():
    assert list(shard_batch(8, 1, 2)) == [4, 5, 6, 7]
    assert [len(shard_batch(64, r, 8)) for r in range(8)] == [8] * 8
    with pytest.raises(ValueError):
        shard_batch(63, 0, 8)


def test_single_process_reducer_is_identity():
    net = _net()
    red = GradAllReducer(net.parameters(), bucket_bytes=256)
    x = torch.randn(4, 16)
    red.prepare()
    net[:5](x).sum().backward()
    before = [p.grad.clone() if p.grad is not None else None for p in net.parameters()]
    red.finish()
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p.grad, b if b is not None else torch.zeros_like(p))


class _TwoHeads(nn.Module):
    """Trunk with two heads: video ranks use one, image ranks the other (losses.VideoImageLoss(is_video=...))."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.trunk = nn.Sequential(nn.Linear(12, 24), nn.GELU(), nn.Linear(24, 24))
        self.head_video = nn.Linear(24, 7)
        self.head_image = nn.Linear(24, 4)


def _hetero_worker(rank, world, port, overlap, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _TwoHeads()
        red = GradAllReducer(net.parameters(), bucket_bytes=64, overlap=overlap)  # tiny buckets: heads in different buckets
        assert len(red.buckets) >= 4
        x = torch.randn(6, 12, generator=torch.Generator().manual_seed(10 + rank))
        for _ in range(2):
            for p in net.parameters():
                p.grad = None
            red.prepare()
            feat = net.trunk(x)
            loss = net.head_video(feat).square().mean() if rank == 0 else net.head_image(feat).square().mean()
            loss.backward()
            red.finish()
        torch.save({n: p.grad.clone() for n, p in net.named_parameters()}, out + f".{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_ranks_with_different_unused_parameters(tmp_path, overlap):
    """ranks whose gradient-less parameter sets differ must still issue the bucket all-reduces in the
    same order (fixed bucket order); the result is the mean over ranks with zeros for the unused heads."""
    out = str(tmp_path / "g.pt")
    mp.spawn(_hetero_worker, args=(2, _free_port(), overlap, out), nprocs=2, join=True)
    got0, got1 = torch.load(out + ".0"), torch.load(out + ".1")
    want = {}
    for rank in range(2):
        net = _TwoHeads()
        x = torch.randn(6, 12, generator=torch.Generator().manual_seed(10 + rank))
        feat = net.trunk(x)
        (net.head_video(feat).square().mean() if rank == 0 else net.head_image(feat).square().mean()).backward()
        for n, p in net.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            want[n] = want.get(n, 0) + g / 2
    for n in want:
        assert torch.allclose(got0[n], want[n], rtol=1e-5, atol=1e-7), n
        assert torch.equal(got0[n], got1[n]), n


def test_finish_without_prepare_raises():
    net = _net()
    red = GradAllReducer(net.parameters(), bucket_bytes=256)
    net[:5](torch.randn(4, 16)).sum().backward()
    with pytest.raises(RuntimeError, match="prepare"):
        red.finish()
