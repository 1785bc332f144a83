"""World-size-2 CPU (gloo) tests of the multi-GPU host logic: the C1 statistics payload / all-gather
/ exact Chan merge, the C2 flat gradient all-reduce, and the sharding identity the design rests on -- per-rank backward of
the GLOBAL alignment loss restricted to the rank's samples, summed over ranks, equals the single-process full-batch
gradient of the reference formulation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _chan_merge(n_a, m_a, M2_a, n_b, m_b, M2_b):
    n = n_a + n_b
    d = m_b - m_a
    return n, m_a + d * (n_b / n), M2_a + M2_b + d * d * (n_a * n_b / n)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vitta_b200 import ops
    from oracle import vitta_oracle as O
    torch.manual_seed(0)
    # the full batch is generated identically on every rank; rank r owns a ragged shard of the samples
    layers = [(6, 40, 8), (6, 40, 16)]           # (samples, tokens per sample, channels)
    full = [torch.randn(n, t, c) * (1.0 + i) + 0.3 * i for i, (n, t, c) in enumerate(layers)]
    split = [0, 4, 6]                             # rank 0: samples 0-3, rank 1: samples 4-5
    mine = [x[split[rank]:split[rank + 1]] for x in full]
    total_C = sum(c for _, _, c in layers)
    # ---- C1: per-rank (mean, M2, n) -> payload -> all-gather -> exact merge
    pay, merged, cnts = ops.stats_payload(total_C, len(layers), "cpu")
    off = 0
    for li, x in enumerate(mine):
        f = x.reshape(-1, x.shape[-1])
        merged[2 * off:2 * (off + f.shape[1]):2] = f.mean(0)
        merged[2 * off + 1:2 * (off + f.shape[1]):2] = ((f - f.mean(0)) ** 2).sum(0)
        cnts[li] = f.shape[0]
        off += f.shape[1]
    means, counts = ops.gather_stats_payload(pay, total_C, dist.group.WORLD)
    assert means.shape == (world, 2 * total_C) and counts.shape == (world, len(layers))
    assert counts.dtype == torch.int32 and counts[:, 0].tolist() == [4 * 40, 2 * 40]
    off = 0
    for li, x in enumerate(full):
        c = x.shape[-1]
        n, m, M2 = 0.0, torch.zeros(c), torch.zeros(c)
        for r in range(world):
            blk = means[r, 2 * off:2 * (off + c)]
            n, m, M2 = _chan_merge(n, m, M2, float(counts[r, li]), blk[0::2], blk[1::2])
        f = x.reshape(-1, c)
        torch.testing.assert_close(m, f.mean(0), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(M2 / n, f.var(0, unbiased=False), rtol=1e-5, atol=1e-6)
        off += c
    # ---- C2: flat gradient all-reduce
    grads = [torch.full((3, 2), float(rank + 1)), torch.arange(5.0) * (rank + 1)]
    outs, flat = ops.allreduce_grads(grads, dist.group.WORLD)
    assert flat.numel() == 11
    torch.testing.assert_close(outs[0], torch.full((3, 2), 3.0))
    torch.testing.assert_close(outs[1], torch.arange(5.0) * 3)
    # ---- sharding identity: sum over ranks of d(global loss)/d(theta) on the local samples == full-batch gradient
    w = torch.randn(8, 8, requires_grad=True)
    src_m, src_v = torch.randn(8) * 0.1, torch.rand(8) + 0.5

    def feature(x):          # a toy "layer": y = x @ w, channels-last tokens (n, t, c) -> the hook's (N, C, T, 1, 1)
        return (x @ w).permute(0, 2, 1)[..., None, None]
    yf = feature(full[0])
    mean, var = O.spatiotemp_stats(yf)
    loss = O.regularization(src_m, 0.1 * mean, src_v, 0.1 * var, "l1_loss")
    g_full, = torch.autograd.grad(loss, w)
    # rank-local: global statistics from the gathered moments, closed-form coefficients
    yl = feature(mine[0])
    with torch.no_grad():
        fl = yl.permute(0, 2, 3, 4, 1).reshape(-1, 8)
        p2, m2, c2 = ops.stats_payload(8, 1, "cpu")
        m2[0::2] = fl.mean(0)
        m2[1::2] = ((fl - fl.mean(0)) ** 2).sum(0)
        c2[0] = fl.shape[0]
        mm, cc = ops.gather_stats_payload(p2, 8, dist.group.WORLD)
        n, gm, gM2 = 0.0, torch.zeros(8), torch.zeros(8)
        for r in range(world):
            n, gm, gM2 = _chan_merge(n, gm, gM2, float(cc[r, 0]), mm[r, 0::2], mm[r, 1::2])
        gvar = gM2 / n
        a = 0.1 * torch.sign(0.1 * gm - src_m) / 8 / n
        b = 0.1 * 2 * torch.sign(0.1 * gvar - src_v) / 8 / n
        gy = (a + b * (fl - gm)).reshape(yl.shape[0], -1, 8).permute(0, 2, 1)[..., None, None]
    g_local, = torch.autograd.grad(yl, w, gy)
    (g_sum,), _ = ops.allreduce_grads([g_local], dist.group.WORLD)
    torch.testing.assert_close(g_sum, g_full, rtol=1e-4, atol=1e-7)
    # ---- driver-level sharding (corpus.basics.tta_standard under torchrun): contiguous video blocks of every GLOBAL
    #      batch, ragged tails split as evenly as possible, accuracy merged over all videos
    from vitta_b200.corpus.basics import merge_meters, shard_batch
    from vitta_b200.utils.utils_ import AverageMeter
    meter = AverageMeter()
    seen = []
    for bz in (8, 5, 2, 3):
        x = torch.arange(bz * 4.0).reshape(bz, 4)
        y = torch.arange(bz)
        xs, ys = shard_batch(x, y, rank, world)
        assert xs.shape[0] == ys.shape[0] and (xs[:, 0] == ys * 4.0).all()          # videos stay with their labels
        assert xs.shape[0] in (bz // world, bz // world + 1) and (rank != 0 or xs.shape[0] == (bz + 1) // world)
        got = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, torch.tensor([xs.shape[0]]))
        assert sum(int(g) for g in got) == bz                                        # every video exactly once
        seen += ys.tolist()
        # "accuracy" of the shard: share of even labels, weighted by the local video count like tta_standard does
        if xs.shape[0]:
            meter.update(100.0 * float((ys % 2 == 0).float().mean()), xs.shape[0])
    (acc,) = merge_meters([meter], dist.group.WORLD)
    want = 100.0 * sum(1 for bz in (8, 5, 2, 3) for v in range(bz) if v % 2 == 0) / 18
    assert abs(acc - want) < 1e-5, (acc, want)       # shard accuracies are float32 means
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_statistics_and_gradients(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


class _LabelDataset(torch.utils.data.Dataset):
    """Item i = a tensor filled with its label i (so a shard shows which videos it holds)."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return torch.full((3, 2, 2), float(i)), i


def _driver_worker(rank, world, port, out_dir):
    """tta_standard's loop under a 2-rank group with the adapter replaced by a recorder: every rank must see exactly its
    block of every global batch (adaptation AND evaluation loader) and return the accuracy over ALL videos."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vitta_b200.corpus import basics
    from vitta_b200.utils.opts import default_args
    seen = {"adapt": [], "eval": []}

    class Recorder:
        def __init__(self, model_origin, args, stats=None, process_group=None):
            assert process_group is dist.group.WORLD

        def adapt(self, input, target=None, criterion=None, global_videos=None):
            seen["adapt"].append(input[:, 0, 0, 0].tolist())
            seen.setdefault("global", []).append(global_videos)
            return {"loss_ce": None, "loss_reg": torch.tensor(1.0), "loss_consis": None}

        def adapt_idle(self, global_videos):
            seen.setdefault("idle", []).append(global_videos)

        def evaluate(self, input):
            vid = input[:, 0, 0, 0].long()
            seen["eval"].append(vid.tolist())
            logits = torch.zeros(len(vid), 16)
            logits[torch.arange(len(vid)), torch.where(vid % 3 == 0, vid, vid + 1) % 16] = 1.0      # right iff vid % 3 == 0
            return logits

        def hooks_off(self):
            pass

        def hooks_on(self):
            pass
    basics.OnlineAdapter = Recorder
    n_videos = 11                                       # batches of 4, 4, 3: the last one is ragged
    args = default_args(arch="tanet", batch_size=4, workers=0, num_classes=16, verbose=False, stat_reg="BNS")
    args.process_group = dist.group.WORLD
    args.dataset_factory = lambda a, split, kind: _LabelDataset(n_videos)
    (acc,) = basics.tta_standard(torch.nn.Linear(2, 2), None, args=args)
    want_blocks = {0: [[0.0, 1.0], [4.0, 5.0], [8.0, 9.0]], 1: [[2.0, 3.0], [6.0, 7.0], [10.0]]}[rank]
    assert seen["adapt"] == want_blocks and seen["eval"] == [[int(v) for v in b] for b in want_blocks], seen
    assert seen["global"] == [4, 4, 3] and "idle" not in seen          # every rank is told the GLOBAL batch size
    want_acc = 100.0 * sum(1 for v in range(n_videos) if v % 3 == 0) / n_videos
    assert abs(acc - want_acc) < 1e-4, (acc, want_acc)
    # ragged tail with FEWER videos than ranks (used to raise after all the work was done): 5 videos in
    # batches of 4 -> the last batch has 1 video; rank 1 idles through the step's collectives and skips the meters
    for k in ("adapt", "eval", "global", "idle"):
        seen[k] = []
    args.dataset_factory = lambda a, split, kind: _LabelDataset(5)
    (acc,) = basics.tta_standard(torch.nn.Linear(2, 2), None, args=args)
    if rank == 0:
        assert seen["adapt"] == [[0.0, 1.0], [4.0]] and seen["global"] == [4, 1] and seen["idle"] == [], seen
    else:
        assert seen["adapt"] == [[2.0, 3.0]] and seen["idle"] == [1] and seen["eval"] == [[2, 3]], seen
    assert abs(acc - 100.0 * 2 / 5) < 1e-4, acc
    # a batch size below the world size is refused up front, before any adaptation work
    args.batch_size = 1
    try:
        basics.tta_standard(torch.nn.Linear(2, 2), None, args=args)
        raise AssertionError("batch_size < world must be refused")
    except ValueError as e:
        assert "cannot be sharded" in str(e)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_driver_shards_every_batch_and_merges_accuracy(tmp_path):
    port = _free_port()
    mp.spawn(_driver_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def _bucket_worker(rank, world, port, out_dir):
    """GradBucketer (collective C2 overlapped with the backward): bucketed asynchronous all-reduces started from
    post-accumulate-grad hooks must deliver exactly the summed gradients of a flat all-reduce, leave p.grad alone,
    keep their flat-buffer addresses across steps, and stand down cleanly when disabled or when the live set changes."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vitta_b200 import ops
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(16, 16) for _ in range(6)])
    params = list(net.parameters())
    bk = ops.GradBucketer(params, dist.group.WORLD, n_buckets=3)
    assert len(bk.buckets) == 3 and bk.buckets[0][0] is params[-1]          # reverse registration order
    assert sum(len(b) for b in bk.buckets) == len(params)
    ptrs = {p: bk.views[p].data_ptr() for p in params}
    for step in range(3):
        for p in params:
            p.grad = None
        x = torch.randn(4, 16, generator=torch.Generator().manual_seed(10 * step + rank))
        net(x).square().sum().backward()
        local = [p.grad.clone() for p in params]
        views = bk.finish(params)
        assert views is not None
        for p, g in zip(params, local):
            want = g.clone()
            dist.all_reduce(want)
            assert torch.equal(p.grad, g)                                   # p.grad untouched
            assert torch.allclose(views[p], want, rtol=0, atol=0)           # same sum (2 ranks: order-independent)
            assert views[p].data_ptr() == ptrs[p]                           # stable addresses
    # disabled (ragged step): hooks stand down, finish() says "flat path"
    bk.enabled = False
    for p in params:
        p.grad = None
    net(torch.randn(4, 16)).sum().backward()
    assert bk.finish(params) is None
    bk.enabled = True
    # a parameter without gradient this step: its bucket never completes -> None, and no collective is left dangling
    for p in params:
        p.grad = None
    net[3:](torch.randn(4, 16)).sum().backward()
    live = [p for p in params if p.grad is not None]
    assert len(live) == 6 and bk.finish(live) is None
    dist.barrier()
    bk.close()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_bucket_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
