"""GPU tests of the training-step plumbing: the CUDA-graph replay of a whole step (GraphedTrainStep) must follow the
eager step (forward, CE loss, backward, clip, fused AdamW) bit for bit when nothing random is left in the model, and
keep drawing fresh DropPath / dropout masks when there is."""
import copy

import pytest
import torch
import torch.nn as nn

import svit_b200
from svit_b200 import ops
from svit_b200.config import tiny_cfg
from svit_b200.optim import construct_optimizer
from tests.golden.recipe import synth_input

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tiny(seed=0, stochastic=False):
    cfg = tiny_cfg()
    torch.manual_seed(seed)
    m = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).to(DEV).train()
    if not stochastic:
        for mod in m.modules():
            if isinstance(mod, svit_b200.DropPath):
                mod.drop_prob = 0.0
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
    return cfg, m


def _eager_step(m, opt, clip, labels, max_norm):
    for p in m.parameters():
        p.grad = None
    _, extra = m([clip])
    loss = torch.nn.functional.cross_entropy(extra["logits"].float(), labels)
    loss.backward()
    opt.step(max_norm=max_norm)
    return loss.detach()


def test_graphed_train_step_matches_eager_steps():
    """Same start, same clips: the replayed graph and the eager loop must follow the same trajectory.  The backward
    uses fp32 atomics (split-K weight gradients, pooling-weight gradients), so two runs differ in the last bits and
    Adam turns the sign of a noise-level gradient into a full +-lr move of that element: the comparison is therefore on
    the loss values (2 %) and on the direction / size of the total parameter movement, not element by element."""
    cfg, m_e = _tiny()
    m_g = copy.deepcopy(m_e)
    init = [p.detach().clone() for p in m_e.parameters()]
    clips = [synth_input(f"train.clip{i}", (2, 3, 4, 32, 32), 7 + i).to(DEV).bfloat16() for i in range(3)]
    labels = [torch.tensor([1, 3], device=DEV), torch.tensor([0, 9], device=DEV), torch.tensor([4, 4], device=DEV)]
    opt_e, opt_g = construct_optimizer(m_e, cfg), construct_optimizer(m_g, cfg)
    for o in (opt_e, opt_g):
        for g in o.param_groups:
            g["lr"] = 2e-3
    warm = 2
    step = svit_b200.GraphedTrainStep(m_g, opt_g, clips[0], labels[0], max_norm=1.0, warmup=warm)
    # the constructor ran `warm` eager steps on (clips[0], labels[0]); do the same on the eager twin
    for _ in range(warm):
        _eager_step(m_e, opt_e, clips[0], labels[0], 1.0)
    losses_e, losses_g = [], []
    for i in range(6):
        if i == 3:  # the learning-rate schedule keeps working between replays
            for o in (opt_e, opt_g):
                for g in o.param_groups:
                    g["lr"] = 5e-4
        losses_e.append(float(_eager_step(m_e, opt_e, clips[i % 3], labels[i % 3], 1.0)))
        losses_g.append(float(step(clips[i % 3], labels[i % 3])))
    assert losses_e == pytest.approx(losses_g, rel=2e-2), (losses_e, losses_g)
    assert losses_e[3] < losses_e[0]  # clip 0 again after three more steps: the optimizer is learning
    assert opt_e._step == opt_g._step == warm + 6
    de = torch.cat([(p.detach() - q).flatten() for p, q in zip(m_e.parameters(), init)])
    dg = torch.cat([(p.detach() - q).flatten() for p, q in zip(m_g.parameters(), init)])
    assert de.norm() > 0 and abs(float(dg.norm() / de.norm()) - 1.0) < 0.05
    assert float(torch.dot(de, dg) / (de.norm() * dg.norm())) > 0.9
    # an eager forward after the replays must see the updated weights (weight caches are invalidated per replay)
    m_g.eval()
    with torch.no_grad():
        _, ex = m_g([clips[0]])
    got = float(torch.nn.functional.cross_entropy(ex["logits"].float(), labels[0]))
    stale = copy.deepcopy(m_g)
    for p, q in zip(stale.parameters(), init):
        p.data.copy_(q)
    with torch.no_grad():
        _, ex0 = stale.eval()([clips[0]])
    first = float(torch.nn.functional.cross_entropy(ex0["logits"].float(), labels[0]))
    assert got < first  # trained weights, not the capture-time copies


def test_graphed_train_step_draws_fresh_droppath_masks():
    """DropPath / head dropout inside the graph use torch's graph-safe generator: replays on the same input give
    different losses (fresh masks), as the eager loop does."""
    cfg, m = _tiny(stochastic=True)
    for mod in m.modules():
        if isinstance(mod, svit_b200.DropPath):
            mod.drop_prob = 0.5
    opt = construct_optimizer(m, cfg)
    for g in opt.param_groups:
        g["lr"] = 0.0  # freeze the weights: only the masks change between replays
        g["weight_decay"] = 0.0
    clip = synth_input("train.clip0", (2, 3, 4, 32, 32), 7).to(DEV).bfloat16()
    labels = torch.tensor([1, 3], device=DEV)
    step = svit_b200.GraphedTrainStep(m, opt, clip, labels, max_norm=None, warmup=1)
    vals = {round(float(step(clip, labels)), 6) for _ in range(8)}
    assert len(vals) > 1


def test_graphed_forward_lanes_match_single_lane():
    """GraphedForward(lanes=n) captures n independent sub-batch forwards on separate streams; the clips do not interact,
    so every output must equal the single-lane graph bit for bit (same kernels per clip, different co-scheduling)."""
    cfg, m = _tiny()
    m.eval()
    clip = synth_input("lanes.clip", (6, 3, 4, 32, 32), 9).to(DEV).bfloat16()
    g1 = svit_b200.GraphedForward(m, clip)
    p1, e1 = g1(clip)
    p1, e1 = p1.clone(), {k: v.clone() for k, v in e1.items() if torch.is_tensor(v)}
    for lanes in (2, 3):
        g = svit_b200.GraphedForward(m, clip, lanes=lanes)
        for _ in range(2):
            p, e = g(clip)
        torch.cuda.synchronize()
        assert p.shape == p1.shape and torch.equal(p, p1), lanes
        for k, v in e1.items():
            assert e[k].shape == v.shape and torch.equal(e[k], v), (lanes, k)


def test_rel_pos_cache_follows_the_parameters():
    """Without autograd every MultiScaleAttention caches its rel-pos tables (gathered R, concatenated bf16 table, index
    tables) per (grid, dtype, device, parameter versions).  An in-place parameter update must invalidate the entry: the
    eager forward and a GraphedForward (which re-captures on a version change) have to see the new tables, and the cached
    forward must equal the forward with emptied caches bit for bit."""
    cfg, m = _tiny()
    m.eval()
    clip = synth_input("relcache.clip", (2, 3, 4, 32, 32), 11).to(DEV).bfloat16()
    with torch.no_grad():
        a1 = m([clip])[1]["logits"].clone()   # fills the caches
        a2 = m([clip])[1]["logits"].clone()   # served from them
    attn = [mod for mod in m.modules() if isinstance(mod, svit_b200.MultiScaleAttention)]
    assert torch.equal(a1, a2)
    assert all(len(mod._rel_cache) == 1 for mod in attn)
    g = svit_b200.GraphedForward(m, clip)
    assert torch.equal(g(clip)[1]["logits"], a1)
    with torch.no_grad():
        for mod in attn:
            mod.rel_pos_h.add_(0.5 * torch.randn_like(mod.rel_pos_h))
            mod.rel_pos_t.mul_(-1.0)
        b1 = m([clip])[1]["logits"].clone()
        for mod in attn:
            mod._rel_cache.clear()
        ref2 = m([clip])[1]["logits"].clone()  # tables rebuilt from the parameters
    assert not torch.equal(b1, a1), "the update of the tables must change the logits"
    assert torch.equal(b1, ref2)
    assert all(len(mod._rel_cache) == 1 for mod in attn), "stale entries are dropped"
    assert torch.equal(g(clip)[1]["logits"], b1), "the graph re-captures on a parameter version change"


def test_graphed_forward_on_uint8_frames_equals_eager():
    """The end-to-end leg of bench.py replays a graph captured on the decoded uint8 frames [B, T, H, W, 3] (normalisation fused
    into the stem's cell-layout kernel): same result as the eager uint8 forward and as the forward of the pre-normalised
    bf16 clip."""
    cfg, m = _tiny()
    m.eval()
    gen = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (2, 4, 32, 32, 3), generator=gen, dtype=torch.uint8).to(DEV)
    with torch.no_grad():
        eager = m([frames])[1]["logits"].clone()
        clip = ops.normalize_u8(frames, cfg.DATA.MEAN, cfg.DATA.STD, torch.bfloat16)
        from_clip = m([clip])[1]["logits"].clone()
    g = svit_b200.GraphedForward(m, frames)
    frames2 = torch.randint(0, 256, (2, 4, 32, 32, 3), generator=gen, dtype=torch.uint8).to(DEV)
    g(frames2)
    out = g(frames)[1]["logits"]
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    assert torch.equal(eager, from_clip)
