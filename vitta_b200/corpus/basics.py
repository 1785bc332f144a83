"""Adaptation driver -- mirror of the hot-path parts of the reference's ``corpus/basics.py``:
``tta_standard`` (:403-747), ``compute_statistics`` (:220-307), ``validate`` (:149-217, source-only evaluation)
and ``get_model`` (:1447-1493).  The loop body lives in :class:`OnlineAdapter` so that ``bench.py`` and the
tests can drive single steps; ``tta_standard`` wraps it in the reference's loader loop.

Out of scope: the decord/PIL data pipeline.  ``get_dataset_tanet`` /
``get_dataset_videoswin`` therefore return synthetic tensor datasets unless ``args.dataset_factory`` is set to
a callable ``(args, split, dataset_type) -> torch Dataset`` yielding the reference's loader layouts.
"""
import copy as cp
import os.path as osp
import time

import numpy as np
import torch
import torch.nn as nn

from .. import ops, synth
from ..utils.BNS_utils import BNFeatureHook, choose_layers, collect_bn_params, freeze_except_bn
from ..utils.norm_stats_utils import CombineNormStatsRegHook_onereg, ComputeNormStatsHook
from ..utils.pred_consistency_utils import compute_pred_consis
from ..utils.utils_ import AverageMeter, accuracy

CANDIDATE_BN_LAYERS = [nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d]


# ----------------------------------------------------------------------------------------------
# datasets (synthetic stand-ins for the reference's video loaders)
# ----------------------------------------------------------------------------------------------
class SyntheticVideoDataset(torch.utils.data.Dataset):
    """Loader-layout tensors: TANet (M*T*3, H, W) per item, Swin (M, 3, T, H, W) per item."""

    def __init__(self, arch, n_items, n_views, clip_len, size, num_classes, seed, tag):
        v = synth.synth_video(n_items, n_views, clip_len, size, seed=seed, gauss_sigma=0.38, tag=tag)
        self.x = synth.tanet_loader_tensor(v) if arch == 'tanet' else synth.swin_loader_tensor(v)
        self.y = synth.synth_labels(n_items, num_classes, seed)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i]


def _dataset(args, split, dataset_type):
    factory = getattr(args, 'dataset_factory', None)
    if factory is not None:
        return factory(args, split, dataset_type)
    views = args.n_augmented_views if (dataset_type == 'tta' and args.if_sample_tta_aug_views) else 1
    return SyntheticVideoDataset(args.arch, getattr(args, 'synthetic_items', 4 * args.batch_size), views,
                                 args.clip_length, args.input_size, args.num_classes,
                                 seed=getattr(args, 'synthetic_seed', 0), tag='tta')


def _loader(dataset, args):
    """The reference's DataLoader settings (shuffle off, pinned host batches); datasets whose items already live on the
    device (``on_device``: vitta_b200.corpus.views.Decoded*VideoDataset) are iterated in-process and not pinned."""
    on_device = getattr(dataset, 'on_device', False)
    return torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, shuffle=False,
                                       num_workers=0 if on_device else args.workers, pin_memory=not on_device)


def get_dataset_tanet(args, split='train', dataset_type=None):
    if split == 'train':      # reference corpus/basics.py:1225-1226 (and its default): adaptation uses split='val' only
        raise NotImplementedError('Training dataset processing for TANet to be added!')
    return _dataset(args, split, dataset_type)


def get_dataset_videoswin(args, split='train', dataset_type=None):
    if split == 'train':      # reference corpus/basics.py:1194-1195
        raise NotImplementedError('Training dataset processing for Video Swin Transformer to be added!')
    return _dataset(args, split, dataset_type)


# ----------------------------------------------------------------------------------------------
# video sharding across ranks (the reference itself is single-process)
# ----------------------------------------------------------------------------------------------
def shard_batch(input, target, rank, world):
    """Rank r's videos of one GLOBAL loader batch: every rank iterates the same loader (same order, same batches as the
    single-process run) and keeps a contiguous block of ``batch / world`` videos with all their views, so that the merged
    statistics (collective C1) and summed gradients (C2) are those of the reference's full batch.  A ragged last batch is
    split as evenly as possible (the first ``batch % world`` ranks take one more video); a rank may get none."""
    bz = input.shape[0]
    base, rem = divmod(bz, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return input[lo:hi], target[lo:hi]


def merge_meters(meters, process_group):
    """Global value of weighted-average meters whose per-rank weights are the local video counts: one all-reduce of
    (sum, count) pairs.  Returns the list of global averages (identical on every rank)."""
    import torch.distributed as dist
    dev = 'cuda' if dist.get_backend(process_group) == 'nccl' else 'cpu'
    buf = torch.tensor([[m.sum, m.count] for m in meters], dtype=torch.float64, device=dev)
    dist.all_reduce(buf, group=process_group)
    return [float(s / c) if c > 0 else 0.0 for s, c in buf.tolist()]


def get_model(args, num_classes, logger=None):
    """reference :1447-1493 (only the two architectures ``--arch`` can select, utils/opts.py:43)."""
    if args.arch == 'tanet':
        from ..models.tanet_models.tanet import TSN
        return TSN(num_classes, args.clip_length, args.modality, base_model='resnet50', consensus_type='avg',
                   img_feature_dim=args.img_feature_dim, tam=True, non_local=False, partial_bn=args.partial_bn)
    if args.arch == 'videoswintransformer':
        from ..models.videoswintransformer_models.recognizer3d import Recognizer3D
        return Recognizer3D(num_classes=num_classes, patch_size=args.patch_size, window_size=args.window_size,
                            drop_path_rate=args.drop_path_rate)
    raise Exception(f'{args.arch} is not a valid model!')


def load_source_statistics(args):
    """Two pickled object arrays, one (C,) vector per BN2d/3d (TANet) or LN[1:] (Swin) in named_modules()
    order (reference :482-483).  ``args.source_stats`` = (mean_list, var_list) bypasses the files."""
    given = getattr(args, 'source_stats', None)
    if given is not None:
        return list(given[0]), list(given[1])
    return (list(np.load(args.spatiotemp_mean_clean_file, allow_pickle=True)),
            list(np.load(args.spatiotemp_var_clean_file, allow_pickle=True)))


def save_stat_list(path, vectors):
    """The reference np.save()s a ragged python list (:306-307), which numpy>=1.24 refuses; the same file
    format (1-D object array of float32 vectors) is built explicitly here."""
    arr = np.empty(len(vectors), dtype=object)
    for i, v in enumerate(vectors):
        arr[i] = np.asarray(v, dtype=np.float32)
    np.save(path, arr, allow_pickle=True)


def _reshape_input(args, x, n_views_or_clips):
    """TANet: (bz, views*T*3, H, W) -> (bz*views, T, 3, H, W) (reference :618-623); Swin: unchanged."""
    if args.arch == 'tanet':
        bz = x.shape[0]
        x = x.view(-1, 3, x.size(2), x.size(3))
        return x.view(bz * args.test_crops * n_views_or_clips, args.clip_length, 3, x.size(2), x.size(3))
    if args.arch == 'videoswintransformer':
        return x
    raise NotImplementedError(f'Incorrect model type {args.arch}')


# ----------------------------------------------------------------------------------------------
# the loop body
# ----------------------------------------------------------------------------------------------
class _Prefetched:
    """Handle of an in-flight host-to-device input copy (OnlineAdapter.prefetch)."""
    __slots__ = ("buf", "ready", "slot")

    def __init__(self, buf, ready, slot):
        self.buf, self.ready, self.slot = buf, ready, slot


class OnlineAdapter:
    """One model copy + optimiser + alignment hooks: everything ``tta_standard`` sets up when
    ``setup_model_optimizer`` is true (reference :525-601), and its per-batch body (:606-728)."""

    def __init__(self, model_origin, args, stats=None, process_group=None):
        self.args = args
        self.process_group = process_group
        # CUDA-graph replay of the whole adaptation step (args.cuda_graph): the step has static shapes, so after a few
        # eager steps it is captured once and replayed -- ~1.6k kernel launches collapse into one graph launch.
        self._graph = None
        self._graph_key = None
        self._graph_in = None
        self._graph_out = None
        self._eager_steps = 0
        self._side = None
        if args.arch == 'tanet':
            self.n_clips = int(args.sample_style.split("-")[-1])
        else:
            self.n_clips = args.num_clips
        if args.if_sample_tta_aug_views:
            assert self.n_clips == 1
        self.if_pred_consistency = args.if_pred_consistency if args.if_sample_tta_aug_views else False
        if not hasattr(args, 'moving_avg'):
            args.moving_avg = False
        if not hasattr(args, 'momentum_mvg'):
            args.momentum_mvg = 0.1

        self.model = cp.deepcopy(model_origin)
        model = self.model
        mean_list = var_list = None
        if args.stat_reg == 'mean_var':
            assert args.stat_type == ['spatiotemp']
            mean_list, var_list = stats if stats is not None else load_source_statistics(args)
            if args.arch == 'tanet':
                self.chosen_layers = choose_layers(model, CANDIDATE_BN_LAYERS)
                it = iter(range(len(mean_list)))
                new_m, new_v = [], []
                for _, layer in self.chosen_layers:     # None placeholders at BatchNorm1d positions (:488-498)
                    if isinstance(layer, nn.BatchNorm1d):
                        new_m.append(None)
                        new_v.append(None)
                    else:
                        i = next(it)
                        new_m.append(mean_list[i])
                        new_v.append(var_list[i])
                mean_list, var_list = new_m, new_v
            elif args.arch == 'videoswintransformer':
                self.chosen_layers = choose_layers(model, [nn.LayerNorm])[1:]    # skip patch_embed.norm (:541-543)
            assert len(mean_list) == len(self.chosen_layers)

        # optimiser (:547-560)
        if args.update_only_bn_affine:
            kinds = CANDIDATE_BN_LAYERS if args.arch == 'tanet' else [nn.LayerNorm]
            self.model = model = freeze_except_bn(model, bn_condidiate_layers=kinds)
            params, _ = collect_bn_params(model, bn_candidate_layers=kinds)
            self.optimizer = torch.optim.Adam(params, lr=args.lr, betas=(0.9, 0.999), weight_decay=0.)
            self._adam_params = list(params)      # multi-GPU: their gradients are summed over ranks before Adam.step
        else:
            self.optimizer = ops.FusedSGD(model.parameters(), lr=args.lr, momentum=args.momentum,
                                          weight_decay=args.weight_decay, process_group=process_group)

        # hooks (:564-601)
        from ..utils import norm_stats_utils as nsu
        nsu.set_process_group(process_group)
        nsu.new_align_generation()          # this adapter's hooks share ONE arena, and only with each other
        self.stat_reg_hooks = []
        self.hooked_layers = []
        if args.stat_reg == 'mean_var':
            if isinstance(args.stat_type, str):
                raise NotImplementedError('args.stat_type of str  is deprecated, use list instead.')
            for layer_id, (name, layer) in enumerate(self.chosen_layers):
                if any(block_name in name for block_name in args.chosen_blocks):
                    self.stat_reg_hooks.append(CombineNormStatsRegHook_onereg(
                        layer, clip_len=args.clip_length,
                        spatiotemp_stats_clean_tuple=(mean_list[layer_id], var_list[layer_id]),
                        reg_type=args.reg_type, moving_avg=args.moving_avg, momentum=args.momentum_mvg,
                        stat_type_list=args.stat_type, reduce_dim=args.reduce_dim, before_norm=args.before_norm,
                        if_sample_tta_aug_views=args.if_sample_tta_aug_views,
                        n_augmented_views=args.n_augmented_views))
                    self.hooked_layers.append(layer)
        elif args.stat_reg == 'BNS':
            self.chosen_layers = choose_layers(model, CANDIDATE_BN_LAYERS)
            for name, layer in self.chosen_layers:
                if any(block_name in name for block_name in args.chosen_blocks):
                    self.stat_reg_hooks.append(BNFeatureHook(layer, reg_type=args.reg_type,
                                                             running_manner=args.running_manner,
                                                             use_src_stat_in_reg=args.use_src_stat_in_reg,
                                                             momentum=args.momentum_bns))
                    self.hooked_layers.append(layer)
        else:
            raise Exception(f'undefined regularization type {args.stat_reg}')
        self._hooks_on = True

    # -- :606-677 -------------------------------------------------------------------------------
    def _arenas(self):
        seen = []
        for h in self.stat_reg_hooks:
            a = getattr(h, '_arena', None)
            if a is not None and all(a is not b for b in seen):
                seen.append(a)
        return seen

    # -- host -> device input prefetch ----------------------------------------------------------------------------------
    def prefetch(self, host_input):
        """Start the host-to-device copy of a (pinned) loader batch on a copy stream and return a handle that
        :meth:`adapt` / :meth:`evaluate` accept in place of the tensor.  Issued for batch i+1 before ``adapt`` of batch
        i, the 77 MB copy of the benchmark batch runs under the previous step instead of in front of its own (two
        staging buffers; a buffer is only overwritten after the step that read it has consumed it)."""
        dev = next(self.model.parameters()).device
        if not hasattr(self, '_pf'):
            self._pf = {'stream': torch.cuda.Stream(device=dev), 'bufs': [None, None], 'free': [None, None], 'i': 0}
        pf = self._pf
        k = pf['i'] % 2
        pf['i'] += 1
        buf = pf['bufs'][k]
        if buf is None or buf.shape != host_input.shape or buf.dtype != host_input.dtype:
            buf = pf['bufs'][k] = torch.empty(host_input.shape, dtype=host_input.dtype, device=dev)
        if pf['free'][k] is not None:
            pf['stream'].wait_event(pf['free'][k])        # the step that used this buffer last has read it
        with torch.cuda.stream(pf['stream']):
            buf.copy_(host_input, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(pf['stream'])
        return _Prefetched(buf, ready, k)

    def _take(self, input):
        """Resolve a prefetch handle: make the compute stream wait for the copy, and note when the buffer is free again."""
        if not isinstance(input, _Prefetched):
            return input, None
        torch.cuda.current_stream().wait_event(input.ready)
        return input.buf, input.slot

    def _release(self, slot):
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._pf['free'][slot] = ev

    def adapt(self, input, target=None, criterion=None, global_videos=None):
        input, slot = self._take(input)
        try:
            return self._adapt(input, target, criterion, global_videos)
        finally:
            self._release(slot)

    def _adapt(self, input, target=None, criterion=None, global_videos=None):
        """One adaptation step.  With ``args.cuda_graph`` (and no label-dependent logging) the step is captured into a
        CUDA graph after 3 eager steps and replayed afterwards; results are identical (same kernels, same order).
        ``global_videos`` (multi-GPU): videos of the whole loader batch this shard was cut from; defaults to
        world x local.  A value below the world size marks a ragged step in which some ranks call :meth:`adapt_idle`."""
        args = self.args
        self._ragged = False
        if self.process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(self.process_group)
            gv = input.shape[0] * world if global_videos is None else int(global_videos)
            n_views = args.n_augmented_views if args.if_sample_tta_aug_views else self.n_clips
            per_video = n_views * (args.test_crops if args.arch == 'tanet' else 1)
            for a in self._arenas():
                a.global_clips = gv * per_video
            self._ragged = gv < world
        graphable = (getattr(args, 'cuda_graph', False) and criterion is None and input.is_cuda and not self._ragged
                     and (self.process_group is None or getattr(args, 'cuda_graph_collectives', False))
                     and getattr(args, 'moving_avg', False)
                     and not args.update_only_bn_affine and args.n_gradient_steps == 1)
        if not graphable:
            return self._adapt_eager(input, target, criterion)
        key = (tuple(input.shape), input.dtype)
        if self._graph is not None and self._graph_key == key:
            return self._replay(input)
        if self._eager_steps < 3 or self._graph_key not in (None, key):
            # Warm-up steps run on a SIDE stream (PyTorch's rule for whole-step capture): autograd's AccumulateGrad nodes
            # remember the stream they were created on, and nodes created on the legacy default stream would make the
            # capture depend on it.
            self._eager_steps += 1
            if self._side is None:
                self._side = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                out = self._adapt_eager(input, target, criterion)
            cur.wait_stream(self._side)
            return out
        from .. import _lib
        self._graph_in = input.clone()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count
        try:
            with torch.cuda.graph(g):
                out = self._adapt_eager(self._graph_in, None, None)
        except _lib.VittaError:
            raise                                         # our own kernels refusing a call is never a capture problem
        except RuntimeError as e:
            # e.g. a collective the installed NCCL cannot capture at this world size: keep adapting with eager launches
            # (same kernels, same order, same results -- only the launch overhead returns) and say so loudly.
            import sys
            sys.stderr.write("vitta_b200: CUDA-graph capture of the adaptation step failed (%s); continuing with eager "
                             "launches\n" % (str(e).splitlines()[0] if str(e) else type(e).__name__))
            args.cuda_graph = False
            self._graph = self._graph_in = None
            _lib.launch_count = l0
            torch.cuda.synchronize()
            return self._adapt_eager(input, target, criterion)
        self._graph_launches = _lib.launch_count - l0     # kernels of ours inside one replay
        _lib.launch_count = l0
        self._graph, self._graph_key, self._graph_out = g, key, out
        return self._replay(input)     # capture does not execute: replay once to actually take this step

    def _replay(self, input):
        from .. import _lib
        if input.data_ptr() != self._graph_in.data_ptr():
            self._graph_in.copy_(input, non_blocking=True)
        self._graph.replay()
        _lib.launch_count += self._graph_launches
        # the replayed step changed the weights behind the split cache's back -- and, being a capture of FusedSGD.step,
        # refreshed the registered splits in place: invalidate everything, then re-validate exactly those
        ops.bump_weight_epoch()
        if isinstance(self.optimizer, ops.FusedSGD):
            ops.refresh_weight_splits(self.optimizer.params, launch=False)
        return dict(self._graph_out)

    def _adapt_eager(self, input, target=None, criterion=None):
        args, model = self.args, self.model
        if not self._hooks_on:
            self.hooks_on()
        model.train()
        if args.fix_BNS:
            for m in model.modules():
                if isinstance(m, tuple(CANDIDATE_BN_LAYERS)):
                    m.eval()
        actual_bz = input.shape[0]
        n_views = args.n_augmented_views if args.if_sample_tta_aug_views else self.n_clips
        x = _reshape_input(args, input, n_views)
        loss_consis = None
        loss_ce = None
        for _ in range(args.n_gradient_steps):
            ops.reset_amax_pool()           # one fill for all operand-range scalars of this pass
            if args.arch == 'tanet':
                output = model(x)
                if args.if_sample_tta_aug_views:
                    output = output.reshape(actual_bz, args.test_crops * n_views, -1)
                    if self.if_pred_consistency:
                        loss_consis = compute_pred_consis(output)
                    output = output.mean(1)
                else:
                    output = output.reshape(actual_bz, args.test_crops * n_views, -1).mean(1)
            else:
                output, view_cls_score = model(x)
                if args.if_sample_tta_aug_views and self.if_pred_consistency:
                    loss_consis = compute_pred_consis(view_cls_score)
            if criterion is not None and target is not None:
                loss_ce = criterion(output.detach(), target)        # logging only (:657)
            loss_reg = self._total_alignment_loss()
            if loss_reg is None:
                loss_reg = torch.zeros((), dtype=torch.float32, device=x.device)
                for hook in self.stat_reg_hooks:
                    loss_reg = loss_reg + hook.r_feature
            if self.if_pred_consistency:
                loss = args.lambda_feature_reg * loss_reg + args.lambda_pred_consis * loss_consis
            else:
                loss = loss_reg                                        # :667: lambda_feature_reg not applied
            self.optimizer.zero_grad()
            if isinstance(self.optimizer, ops.FusedSGD):
                self.optimizer.set_overlap(not getattr(self, '_ragged', False))
            loss.backward()
            if getattr(self, '_ragged', False):
                self._live_mask = self._exchange_live_mask()     # idle ranks are waiting for it (adapt_idle)
            self._sync_foreign_grads()
            self.optimizer.step()
        return {'output': output.detach(), 'loss_reg': loss_reg.detach(), 'loss': loss.detach(),
                'loss_consis': None if loss_consis is None else loss_consis.detach(), 'loss_ce': loss_ce}

    def _total_alignment_loss(self):
        """sum_h hook.r_feature in one autograd node when every contributing hook lives in this adapter's single arena
        (the BatchNorm1d hooks contribute exact zeros, reference utils/norm_stats_utils.py:158-183)."""
        arenas = self._arenas()
        if len(arenas) != 1:
            return None
        layers = []
        for h in self.stat_reg_hooks:
            ly = getattr(h, '_layer', None)
            if ly is None:
                if not getattr(h, '_is_bn1d', False):
                    return None
                continue
            if ly.geom_key is None or not ly.fired:
                return None
            layers.append(ly)
        return arenas[0].total_loss(layers) if layers else None

    def _exchange_live_mask(self):
        if isinstance(self.optimizer, ops.FusedSGD):
            return self.optimizer.exchange_live_mask()
        import torch.distributed as dist
        dev = self._adam_params[0].device
        mask = torch.tensor([1 if p.grad is not None else 0 for p in self._adam_params], dtype=torch.uint8, device=dev)
        dist.broadcast(mask, src=dist.get_global_rank(self.process_group, 0), group=self.process_group)
        return [bool(v) for v in mask.tolist()]

    def adapt_idle(self, global_videos):
        """The step of a rank that holds NO video of a ragged global batch (fewer videos than ranks): zero-count
        statistics into collective C1, zero gradients into C2, and the same optimiser update as everyone else --
        otherwise meters and weights would diverge between ranks.  Same collective order as
        :meth:`_adapt_eager`: C1 (one per arena), live-mask broadcast, C2."""
        args = self.args
        if self.process_group is None:
            raise RuntimeError("adapt_idle: only meaningful under a process group")
        dev = next(self.model.parameters()).device
        n_views = args.n_augmented_views if args.if_sample_tta_aug_views else self.n_clips
        per_video = n_views * (args.test_crops if args.arch == 'tanet' else 1)
        for _ in range(args.n_gradient_steps):
            for a in self._arenas():
                a.global_clips = int(global_videos) * per_video
                a.finalize_idle(dev)
            if isinstance(self.optimizer, ops.FusedSGD):
                self.optimizer.set_overlap(False)
                self.optimizer.step_idle()
            else:
                live = self._exchange_live_mask()
                self.optimizer.zero_grad()
                for p, on in zip(self._adam_params, live):
                    p.grad = torch.zeros_like(p) if on else None
                self._sync_foreign_grads()
                self.optimizer.step()
                self.optimizer.zero_grad()

    def _sync_foreign_grads(self):
        """Collective C2 for the optimisers that are not ours: FusedSGD all-reduces inside step(); torch's Adam on the
        norm-affine parameters (--update_only_bn_affine, reference :547-557) knows nothing about ranks, so the summed
        full-batch gradient is formed here -- every rank then takes the single-process update."""
        if self.process_group is None or isinstance(self.optimizer, ops.FusedSGD):
            return
        live = [p for p in self._adam_params if p.grad is not None]
        if not live:
            return
        summed, _ = ops.allreduce_grads([p.grad for p in live], self.process_group)
        for p, g in zip(live, summed):
            p.grad.copy_(g)

    def hooks_off(self):
        for h in self.stat_reg_hooks:        # :682-684
            h.close()
        self._hooks_on = False

    def hooks_on(self):
        for h, layer in zip(self.stat_reg_hooks, self.hooked_layers):    # :721-727
            h.add_hook_back(layer)
        self._hooks_on = True

    @torch.no_grad()
    def evaluate(self, input):
        """:691-713 -- clean forward on the same videos with hooks removed and model.eval().  With ``args.cuda_graph`` the
        forward (BN-folded convolutions, their two operand-refresh launches included) is captured after one eager pass
        and replayed: in eager mode its ~150 launches are bound by the Python / ctypes launch path, not by the GPU."""
        args, model = self.args, self.model
        input, slot = self._take(input)
        if self._hooks_on:
            self.hooks_off()
        model.eval()
        try:
            if not (getattr(args, 'cuda_graph', False) and input.is_cuda and self.process_group is None
                    and not getattr(self, '_eval_graph_failed', False)):
                return self._evaluate_eager(input)
            key = (tuple(input.shape), input.dtype)
            eg = getattr(self, '_eval_graph', None)
            if eg is not None and eg['key'] == key:
                if input.data_ptr() != eg['in'].data_ptr():
                    eg['in'].copy_(input, non_blocking=True)
                eg['graph'].replay()
                return eg['out'].clone()        # the graph's output buffer is rewritten by the next replay
            if getattr(self, '_eval_eager_key', None) != key:
                # first pass of this shape runs eagerly on a side stream (allocator / cuDNN warm-up before capture)
                self._eval_eager_key = key
                side = self._side if self._side is not None else torch.cuda.Stream()
                self._side = side
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    out = self._evaluate_eager(input)
                cur.wait_stream(side)
                return out
            static_in = input.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    out = self._evaluate_eager(static_in)
            except RuntimeError as e:
                import sys
                sys.stderr.write("vitta_b200: CUDA-graph capture of the evaluation forward failed (%s); staying eager\n"
                                 % (str(e).splitlines()[0] if str(e) else type(e).__name__))
                self._eval_graph = None
                self._eval_graph_failed = True
                torch.cuda.synchronize()
                return self._evaluate_eager(input)
            self._eval_graph = {'key': key, 'graph': g, 'in': static_in, 'out': out}
            g.replay()
            return out.clone()
        finally:
            self._release(slot)

    def _evaluate_eager(self, input):
        args, model = self.args, self.model
        ops.reset_amax_pool()
        x = _reshape_input(args, input, self.n_clips)
        if args.arch == 'tanet':
            out = model(x)
            return out.reshape(input.shape[0], args.test_crops * self.n_clips, -1).mean(1)
        out, _ = model(x)
        return out


def tta_standard(model_origin, criterion, args=None, logger=None, writer=None):
    """Online test-time adaptation over a loader pair (reference :403-747).  ``tta_online``: one adapter for
    the whole stream; ``tta_standard``: a fresh model copy, optimiser and hooks per batch."""
    if args.if_tta_standard == 'tta_standard':
        assert args.momentum_mvg == 1.0
        assert args.n_epoch_adapat == 1
    elif args.if_tta_standard == 'tta_online':
        assert args.momentum_mvg != 1.0
        assert args.n_gradient_steps == 1
        assert args.n_epoch_adapat == 1
    make = get_dataset_tanet if args.arch == 'tanet' else get_dataset_videoswin
    tta_loader = _loader(make(args, split='val', dataset_type='tta'), args)
    eval_loader = _loader(make(args, split='val', dataset_type='eval'), args)
    stats = load_source_statistics(args) if args.stat_reg == 'mean_var' else None
    batch_time, losses_ce, losses_reg, losses_consis = AverageMeter(), AverageMeter(), AverageMeter(), AverageMeter()
    top1, top5 = AverageMeter(), AverageMeter()
    eval_iter = iter(eval_loader)
    device = next(model_origin.parameters()).device
    adapter = None
    end = time.time()
    pg = getattr(args, 'process_group', None)      # set by corpus.main_eval.eval under torchrun: shard the videos
    if pg is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(pg), dist.get_world_size(pg)
        if args.batch_size < world:
            # checked BEFORE any work: every full batch would leave ranks idle (a ragged tail alone is fine, below)
            raise ValueError("--batch_size %d cannot be sharded over %d ranks: launch at most batch_size processes or "
                             "raise the batch size (the global batch is split over the ranks)" % (args.batch_size, world))
    for batch_id, (input, target) in enumerate(tta_loader):
        if args.if_tta_standard == 'tta_standard' or batch_id == 0:
            adapter = OnlineAdapter(model_origin, args, stats, pg)
        global_videos = input.shape[0]
        if pg is not None:
            input, target = shard_batch(input, target, rank, world)
        actual_bz = input.shape[0]
        if actual_bz == 0:
            # ragged tail with fewer videos than ranks: this rank has nothing to forward but still takes part in the
            # step's collectives and applies the common update; it contributes nothing to the accuracy meters
            adapter.adapt_idle(global_videos)
            next(eval_iter)
            continue
        input = input.to(device, non_blocking=True)
        target = target.to(device, non_blocking=True)
        if pg is not None:
            res = adapter.adapt(input, target, criterion, global_videos=global_videos)
        else:
            res = adapter.adapt(input, target, criterion)
        if res['loss_ce'] is not None:
            losses_ce.update(res['loss_ce'].item(), actual_bz)
        losses_reg.update(res['loss_reg'].item(), actual_bz)
        if res['loss_consis'] is not None:
            lc = res['loss_consis']
            if pg is not None:
                # the consistency loss is a SUM over the videos of the batch (utils/pred_consistency_utils.py:15-31): a rank
                # holds the share of its shard (the gradients are summed by C2); the meter shows the whole batch's value,
                # as the single-process run does.  (Ranks without a video are in adapt_idle and do not reach this point:
                # the collective is skipped on ragged steps.)
                if global_videos >= world:
                    lc = lc.clone()
                    dist.all_reduce(lc, group=pg)
            losses_consis.update(lc.item(), actual_bz)
        adapter.hooks_off()
        input, target = next(eval_iter)
        if pg is not None:
            input, target = shard_batch(input, target, rank, world)
        if hasattr(adapter, 'prefetch') and not input.is_cuda and device.type == 'cuda':
            input = adapter.prefetch(input)          # copy stream; evaluate() waits for it
        else:
            input = input.to(device, non_blocking=True)
        target = target.to(device, non_blocking=True)
        output = adapter.evaluate(input)
        prec1, prec5 = accuracy(output.data, target, topk=(1, 5))
        top1.update(prec1.item(), actual_bz)
        top5.update(prec5.item(), actual_bz)
        batch_time.update(time.time() - end)
        end = time.time()
        if args.if_tta_standard == 'tta_online':
            adapter.hooks_on()
        if args.verbose and logger is not None:
            logger.debug(f'TTA Epoch1: [{batch_id}/{len(tta_loader)}]\t'
                         f'Time {batch_time.val:.3f} ({batch_time.avg:.3f})\t'
                         f'Loss reg {losses_reg.val:.4f} ({losses_reg.avg:.4f})\t'
                         f'Loss consis {losses_consis.val:.4f} ({losses_consis.avg:.4f})\t'
                         f'Prec@1 {top1.val:.3f} ({top1.avg:.3f})\tPrec@5 {top5.val:.3f} ({top5.avg:.3f})')
    tta_standard.last_adapter = adapter
    if pg is not None:       # accuracy over ALL videos, identical on every rank
        return [merge_meters([top1], pg)[0]]
    return [top1.avg]


def compute_statistics(model=None, args=None, logger=None, log_time=None):
    """Source-statistics producer (reference :220-307): model.eval(), one ComputeNormStatsHook per BN2d/3d
    (TANet; BN1d too for the 'temp' types) or LN[1:] (Swin); per-batch mean and per-batch *biased variance*
    averaged with weight = batch size; written as two .npy object arrays."""
    if args.arch == 'tanet':
        kinds = CANDIDATE_BN_LAYERS if args.stat_type in ['temp', 'temp_v2'] else [nn.BatchNorm2d, nn.BatchNorm3d]
        chosen = choose_layers(model, kinds)
        n_clips = int(args.sample_style.split("-")[-1])
        loader_ds = get_dataset_tanet(args, split='val', dataset_type='eval')
    elif args.arch == 'videoswintransformer':
        chosen = choose_layers(model, [nn.LayerNorm])[1:]
        n_clips = args.num_clips
        loader_ds = get_dataset_videoswin(args, split='val', dataset_type='eval')
    else:
        raise Exception(f'{args.arch} is not a valid model!')
    hooks = [ComputeNormStatsHook(layer, clip_len=args.clip_length, stat_type=args.stat_type,
                                  before_norm=args.before_norm, batch_size=args.batch_size) for _, layer in chosen]
    loader = _loader(loader_ds, args)
    device = next(model.parameters()).device
    sum_mean = [None] * len(hooks)
    sum_var = [None] * len(hooks)
    count = 0
    model.eval()
    with torch.no_grad():
        for batch_id, (input, _) in enumerate(loader):
            bz = input.shape[0]
            model(_reshape_input(args, input.to(device, non_blocking=True), n_clips))
            for i, h in enumerate(hooks):
                m, v = h.batch_mean * bz, h.batch_var * bz
                sum_mean[i] = m.clone() if sum_mean[i] is None else sum_mean[i] + m
                sum_var[i] = v.clone() if sum_var[i] is None else sum_var[i] + v
            count += bz
    for h in hooks:
        h.close()
    list_mean = [(s / count).cpu().numpy() for s in sum_mean]
    list_var = [(s / count).cpu().numpy() for s in sum_var]
    if getattr(args, 'result_dir', None):
        save_stat_list(osp.join(args.result_dir, f'list_{args.stat_type}_mean_{log_time}.npy'), list_mean)
        save_stat_list(osp.join(args.result_dir, f'list_{args.stat_type}_var_{log_time}.npy'), list_var)
    return list_mean, list_var


@torch.no_grad()
def validate(val_loader, model, criterion, iter=0, epoch=None, args=None, logger=None, writer=None, optimizer=None):
    """Source-only evaluation: model.eval() forward + accuracy (reference :149-217; config 1 of BASELINE.json)."""
    top1, top5, losses = AverageMeter(), AverageMeter(), AverageMeter()
    model.eval()
    device = next(model.parameters()).device
    n_clips = int(args.sample_style.split("-")[-1]) if args.arch == 'tanet' else args.num_clips
    for input, target in val_loader:
        bz = input.shape[0]
        input, target = input.to(device, non_blocking=True), target.to(device, non_blocking=True)
        x = _reshape_input(args, input, n_clips)
        if args.arch == 'tanet':
            output = model(x).reshape(bz, args.test_crops * n_clips, -1).mean(1)
        else:
            output, _ = model(x)
        if criterion is not None:
            losses.update(criterion(output, target).item(), bz)
        p1, p5 = accuracy(output.data, target, topk=(1, 5))
        top1.update(p1.item(), bz)
        top5.update(p5.item(), bz)
    if logger is not None:
        logger.debug(f'Testing Results: Prec@1 {top1.avg:.3f} Prec@5 {top5.avg:.3f} Loss {losses.avg:.5f}')
    return top1.avg
