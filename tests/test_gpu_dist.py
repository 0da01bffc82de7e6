"""2-rank NCCL test of the data-parallel path with the real model: the gradients after the bucketed
all-reduce on two GPUs (half the batch each) equal the single-GPU gradients of the whole batch, fp32 <= 1e-5
(summation-order noise only).  Needs two GPUs: skipped elsewhere (run with `gpurun --gpus 2`; log in profiles/)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(dev):
    import svit_b200
    from svit_b200 import ops
    from svit_b200.config import state_shapes, tiny_cfg
    from tests.golden.recipe import synth_state

    ops.set_impl(gemm=ops.IMPL_SIMT, attn=ops.IMPL_SIMT)
    cfg = tiny_cfg()
    m = svit_b200.SViT(cfg, compute_dtype=torch.float32)
    m.load_state_dict(synth_state(state_shapes(cfg), 300, w_std=0.06))
    m = m.to(dev).train()
    for mod in m.modules():
        if isinstance(mod, svit_b200.DropPath):
            mod.drop_prob = 0.0
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    return cfg, m


def _loss(m, clip, labels):
    logits, extra = m([clip])
    # sum over samples / global batch: rank losses add up to the large-batch mean loss
    return torch.nn.functional.cross_entropy(logits, labels, reduction="sum") / 4 \
        + 0.1 * extra["pred_bboxes"].square().sum() / 4


def _data():
    from tests.golden.recipe import synth_input
    clip = synth_input("dist.clip", (4, 3, 4, 32, 32), 21)
    labels = torch.tensor([1, 3, 0, 7])
    return clip, labels


def _worker(rank, world, port, overlap, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from svit_b200.distributed import GradAllReducer, shard_batch
        cfg, m = _build(dev)
        red = GradAllReducer(m.parameters(), bucket_bytes=256 << 10, overlap=overlap)
        clip, labels = _data()
        idx = shard_batch(4, rank, world)
        c, l = clip[idx.start:idx.stop].to(dev), labels[idx.start:idx.stop].to(dev)
        for _ in range(2):
            for p in m.parameters():
                p.grad = None
            red.prepare()
            # DDP averages over ranks: scale by world so that the average equals the large-batch gradient
            (_loss(m, c, l) * world).backward()
            red.finish()
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({n: p.grad.detach().cpu() for n, p in m.named_parameters()}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_nccl_gradients_equal_large_batch(tmp_path, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), overlap, out), nprocs=2, join=True)
    got = torch.load(out)
    dev = torch.device("cuda", 0)
    cfg, m = _build(dev)
    clip, labels = _data()
    _loss(m, clip.to(dev), labels.to(dev)).backward()
    worst = 0.0
    for n, p in m.named_parameters():
        want = p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(got[n])
        denom = want.abs().max().item()
        if denom < 1e-7:
            assert got[n].abs().max().item() < 1e-6, n
            continue
        err = (got[n] - want).abs().max().item() / denom
        worst = max(worst, err)
        assert err < 1e-5, (n, err)
    print(f"2-rank NCCL vs 1-rank large batch: worst max-rel-err {worst:.2e} over {len(got)} tensors (overlap={overlap})")
