"""Autograd operators over the C-ABI kernels (include/vitta_b200.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every byte of the
hot-path arithmetic below runs in libvitta_b200.so.  Nothing in this file falls back to eager torch
math when the library is missing -- ``_lib.load()`` raises.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

CL = torch.channels_last


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.VittaError("%s: vitta_b200 kernels need CUDA tensors (got %s); there is no CPU path" % (what, t.device))
    if t.dtype != torch.float32:
        raise _lib.VittaError("%s: fp32 only (got %s)" % (what, t.dtype))


def _fused_amax():
    """Producers emit the operand range (max|out|) of their outputs when the consumers run the fp16 split."""
    return _gemm_precision == "f16x3" and _FUSED_AMAX


# Operand-range scalars live in a pre-zeroed pool: one fill per adaptation step instead of one torch.zeros(1) launch per
# producer (~190 per TANet step).  reset_amax_pool() -- called at the start of every adaptation / evaluation pass --
# re-zeroes the pool and starts handing out its slots again; scalars attached to tensors before a reset are disowned
# through the generation number, so a long-lived tensor can never read a recycled slot.
_AMAX_POOL = 4096
_amax_pools = {}      # device -> [pool tensors], [next index]
_amax_gen = 0


def new_amax(dev):
    """A zero-initialised device scalar (shape [1]) for a max|x| accumulation."""
    ent = _amax_pools.get(dev)
    if ent is None:
        ent = _amax_pools[dev] = [[torch.zeros(_AMAX_POOL, dtype=torch.float32, device=dev)], 0]
    pools, i = ent
    if i >= _AMAX_POOL * len(pools):
        pools.append(torch.zeros(_AMAX_POOL, dtype=torch.float32, device=dev))     # rare: a pass with > 4096 producers
    ent[1] = i + 1
    return pools[i // _AMAX_POOL][i % _AMAX_POOL:i % _AMAX_POOL + 1]


def reset_amax_pool():
    global _amax_gen
    _amax_gen += 1
    for ent in _amax_pools.values():
        del ent[0][1:]
        if ent[1]:
            ent[0][0].zero_()
        ent[1] = 0


def _attach_amax(t, am):
    """Remember max|t| on the tensor OBJECT together with its version counter: an attribute cannot outlive the tensor (no
    stale pointer keys), and an in-place update of the tensor invalidates it."""
    t._vitta_amax = (am, t._version, _amax_gen)


def operand_amax(t):
    """max|t| as a device scalar: the value its producer kernel emitted or an earlier call computed (attached to the
    tensor object), else a standalone vitta_amax_f32 pass whose result is attached for the next consumer (a conv input is
    also the operand of its weight gradient)."""
    ent = getattr(t, "_vitta_amax", None)
    if ent is not None and ent[1] == t._version and ent[2] == _amax_gen:
        return ent[0]
    am = amax_f32(t)
    _attach_amax(t, am)
    return am


def as_rows_cl(x):
    """4-D logical (F, C, H, W) tensor in channels_last memory -> (frames, frame_rows, C) geometry."""
    f, c, h, w = x.shape
    return f, h * w, c


def to_cl(x):
    return x.contiguous(memory_format=CL)


# ----------------------------------------------------------------------------------------------
# Statistics arena: one finalize launch (merge + EMA + loss + backward coefficients) for all layers
# ----------------------------------------------------------------------------------------------
class _Layer:
    __slots__ = ("idx", "C", "reg_type", "moving_avg", "momentum", "has_source", "ch_off", "geom_key", "chunking",
                 "part", "count", "meter_count", "fired", "token", "name", "n_batch")


# ----------------------------------------------------------------------------------------------
# multi-GPU host plumbing (device agnostic: exercised on CPU with the gloo backend by tests/test_dist_gloo.py)
# ----------------------------------------------------------------------------------------------
def stats_payload(total_C, n_layers, device):
    """Per-rank payload of collective C1: [2*total_C floats (mean, M2 per channel)] + [n_layers int32 element counts,
    bit-cast to float32 so that ONE all-gather moves both].  Returns (payload, merged view, counts view)."""
    pay = torch.empty(2 * total_C + n_layers, dtype=torch.float32, device=device)
    return pay, pay[:2 * total_C], pay[2 * total_C:].view(torch.int32)


def gather_stats_payload(pay, total_C, group):
    """all-gather the payloads of every rank; returns ((world, 2*total_C) float32 means/M2, (world, n) int32 counts)."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    gath = torch.empty(ws * pay.numel(), dtype=torch.float32, device=pay.device)
    dist.all_gather_into_tensor(gath, pay, group=group)           # collective C1 (NCCL over NVLink on the GPU box)
    g2 = gath.view(ws, -1)
    return g2[:, :2 * total_C].contiguous(), g2[:, 2 * total_C:].contiguous().view(torch.int32)


def allreduce_grads(grads, group):
    """Collective C2: sum the gradients of all ranks through one flat buffer; returns views into it (same order)."""
    import torch.distributed as dist
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    outs, o = [], 0
    for g in grads:
        outs.append(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
    return outs, flat


class GradBucketer:
    """Collective C2 overlapped with the backward pass (device agnostic; exercised on CPU with gloo).

    The parameters that receive gradients are cut into a few buckets in REVERSE registration order -- the order in which
    the backward pass finishes them (layer4 first: 60 % of TANet's gradient bytes are complete when three quarters of
    the backward are still to run).  A post-accumulate-grad hook on every parameter counts its bucket down; the last
    one packs the bucket's gradients into its slice of ONE pre-allocated flat buffer (a single torch.cat(out=...)) and
    starts an asynchronous all-reduce of that slice.  ``finish()`` waits for the collectives and returns views into the
    flat buffer, which the optimiser reads in place (their addresses never change, so the fused-SGD pointer table and a
    captured CUDA graph stay valid).  ``p.grad`` itself is never modified: if a step's live set differs from the one the
    buckets were built for, ``finish()`` returns None and the caller falls back to the unbucketed all-reduce."""

    def __init__(self, live_params, group, n_buckets=4):
        import torch.distributed as dist
        self.group = group
        self.params = list(live_params)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        order = list(reversed(self.params))
        target = (total + n_buckets - 1) // n_buckets
        self.buckets, cur, acc = [], [], 0
        for p in order:
            cur.append(p)
            acc += p.numel()
            if acc >= target and len(self.buckets) < n_buckets - 1:
                self.buckets.append(cur)
                cur, acc = [], 0
        if cur:
            self.buckets.append(cur)
        self.slices, self.views, self.bucket_of, o = [], {}, {}, 0
        for b, ps in enumerate(self.buckets):
            start = o
            for p in ps:
                self.views[p] = self.flat[o:o + p.numel()].view(p.shape)
                self.bucket_of[p] = b
                o += p.numel()
            self.slices.append(self.flat[start:o])
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._dist = dist
        self.enabled = True       # False: hooks stand down and finish() reports "use the flat path" (ragged steps, where
        self.arm()                # a rank without videos cannot run a backward pass and therefore no bucket hooks)

    def arm(self):
        self._pending = [len(ps) for ps in self.buckets]
        self._works = [None] * len(self.buckets)

    def _on_grad(self, p):
        b = self.bucket_of.get(p)
        if not self.enabled or b is None or self._works[b] is not None:
            return
        self._pending[b] -= 1
        if self._pending[b] == 0:
            ps = self.buckets[b]
            if all(q.grad is not None for q in ps):
                torch.cat([q.grad.reshape(-1) for q in ps], out=self.slices[b])
                self._works[b] = self._dist.all_reduce(self.slices[b], group=self.group, async_op=True)

    def finish(self, live):
        """-> {param: summed gradient view} when exactly the bucketed parameters had gradients this step, else None."""
        ok = len(live) == len(self.params) and all(p in self.views for p in live) and all(w is not None for w in self._works)
        for w in self._works:
            if w is not None:
                w.wait()
        views = self.views if ok else None
        self.arm()
        return views

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def pinned_bytes(ctypes_array):
    """uint8 tensor in pinned host memory holding a ctypes structure array (source of non-blocking H2D copies)."""
    raw = bytes(ctypes_array)
    t = torch.empty(len(raw), dtype=torch.uint8).pin_memory()
    t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    return t


class StatsArena:
    """State shared by all alignment hooks attached to one model.

    Forward: each hooked layer writes per-chunk (mean, M2) partials (K1, fused into the norm kernel where
    the model is ours).  ``finalize()`` -- triggered lazily by the first ``r_feature`` read after the
    forward -- merges them (and, with ``process_group`` set, all-gathers per-rank merges: collective C1),
    applies the EMA / average meter, evaluates the L1/MSE/KLD alignment loss per layer and emits the
    per-channel coefficients (a_c, b_c) of dLoss/dy = a_c + b_c*y consumed by the backward kernels.
    """

    def __init__(self, device=None, process_group=None):
        self.device = device
        self.layers = []
        self.frozen = False
        self.finalized = True      # nothing pending
        self.desc_dirty = True
        self.process_group = process_group
        self._desc_dev = None
        self._desc_active = None
        self._base = None
        self.total_C = 0
        self.finalize_calls = 0
        # multi-GPU: clips (videos x views) of the GLOBAL batch of the current step, set by the adaptation driver.  The
        # AverageMeterTensor weights n / (count + n) use the reference's full-batch n (utils/utils_.py:198-202), which a
        # rank cannot derive from its own shard when the batch is ragged.
        self.global_clips = None

    # -- registration ---------------------------------------------------------------------------
    def add_layer(self, C_, src_mean, src_var, reg_type, moving_avg, momentum, name=""):
        ly = _Layer()
        ly.idx = len(self.layers)
        ly.C = int(C_)
        ly.reg_type = _lib.REG_TYPES[reg_type] if src_mean is not None else 0
        ly.moving_avg = bool(moving_avg)
        ly.momentum = float(momentum)
        ly.has_source = src_mean is not None
        ly.ch_off = self.total_C
        ly.geom_key = None
        ly.chunking = None
        ly.part = None
        ly.count = 0
        ly.meter_count = 0
        ly.fired = False
        ly.token = None
        ly.name = name
        ly.n_batch = 1
        self.total_C += ly.C
        self.layers.append(ly)
        self._pending_src = getattr(self, "_pending_src", [])
        self._pending_src.append((src_mean, src_var))
        if self.frozen:
            self._grow()
        return ly

    def _alloc(self, dev):
        n = self.total_C
        z = lambda: torch.zeros(n, dtype=torch.float32, device=dev)
        self.src_mean, self.src_var = z(), torch.ones(n, dtype=torch.float32, device=dev)
        self.ema_mean, self.ema_var = z(), z()
        self.batch_mean, self.batch_var = z(), z()
        self.coef_a, self.coef_b = z(), z()
        for ly, (m, v) in zip(self.layers, self._pending_src):
            if m is not None:
                sl = slice(ly.ch_off, ly.ch_off + ly.C)
                self.src_mean[sl] = torch.as_tensor(m, dtype=torch.float32).to(dev).reshape(-1)
                self.src_var[sl] = torch.as_tensor(v, dtype=torch.float32).to(dev).reshape(-1)
        self.max_C = max(ly.C for ly in self.layers)
        self.loss = torch.zeros(_lib.load().vitta_stats_finalize_loss_floats(len(self.layers), self.max_C),
                                dtype=torch.float32, device=dev)
        self._base = torch.zeros(4, dtype=torch.float32, device=dev)

    def _freeze(self, dev):
        if self.frozen:
            return
        self.device = dev
        self._alloc(dev)
        self.frozen = True
        self.desc_dirty = True

    def _grow(self):
        old = (self.ema_mean, self.ema_var, self.batch_mean, self.batch_var)
        self._alloc(self.device)
        for dst, src in zip((self.ema_mean, self.ema_var, self.batch_mean, self.batch_var), old):
            dst[:src.numel()] = src
        self.desc_dirty = True

    # -- forward side ---------------------------------------------------------------------------
    def partial_buffer(self, ly, O, Cc, I, frames, dev, ln=False):
        """Return the (mean, M2) partial buffer of a layer for this forward, starting a new step if needed.
        ``ln``: the partials come from the LayerNorm kernel (K9), whose chunks are vitta_ln_chunking's."""
        self._freeze(dev)
        if self.finalized:
            self.finalized = False
            for l2 in self.layers:
                l2.fired = False
        key = (O, Cc, I, frames, ln)
        if ly.geom_key != key:
            ch = _lib.ln_chunking(O, Cc, True) if ln else _lib.chunking(O, Cc, I, frames)
            ly.chunking = ch
            ly.geom_key = key
            ly.part = torch.empty(ch.n_entries * Cc * 2, dtype=torch.float32, device=dev)
            ly.count = O * I
            self.desc_dirty = True
        ly.fired = True
        return ly.part

    def set_source(self, ly, mean, var):
        """Replace the target statistics of one layer (BNFeatureHook with use_src_stat_in_reg=False: the target is the
        layer's live running statistics at hook time, utils/BNS_utils.py:61-62)."""
        if not self.frozen:
            raise _lib.VittaError("set_source: the arena has no device buffers yet (call after the layer recorded)")
        sl = slice(ly.ch_off, ly.ch_off + ly.C)
        self.src_mean[sl].copy_(mean.detach().reshape(-1))
        self.src_var[sl].copy_(var.detach().reshape(-1))

    def record(self, ly, feat, O, Cc, I, frames=1):
        """Standalone K1 launch on a feature tensor (hooks on stock torch modules)."""
        part = self.partial_buffer(ly, O, Cc, I, frames, feat.device)
        call("vitta_stats_partial", ptr(feat), O, Cc, I, frames, ptr(part), stream_ptr())

    # -- finalize -------------------------------------------------------------------------------
    def _build_descs(self, active, gathered=None):
        arr = (_lib.VittaLayerDesc * len(active))()
        base = self._base.data_ptr()
        for i, ly in enumerate(active):
            d = arr[i]
            d.C = ly.C
            d.ch_off = ly.ch_off
            d.reg_type = ly.reg_type
            d.has_source = 1 if ly.has_source else 0
            if ly.moving_avg:
                d.w_new, d.w_old = ly.momentum, 1.0 - ly.momentum
            else:
                n = self._meter_n(ly)
                d.w_new = n / float(ly.meter_count + n)
                d.w_old = ly.meter_count / float(ly.meter_count + n)
            if gathered is None:
                ch = ly.chunking
                d.n_entries, d.chunk_rows, d.chunks_per_frame = ch.n_entries, ch.chunk_rows, ch.chunks_per_frame
                d.frame_rows = ch.frame_rows
                off = ly.part.data_ptr() - base
                assert off % 4 == 0
                d.part_off, d.entry_stride = off // 4, 2 * ly.C
                d.cnt_off = d.cnt_stride = 0
            else:
                world, n_active = gathered
                d.n_entries, d.chunk_rows, d.chunks_per_frame, d.frame_rows = world, 0, 1, 0
                d.part_off, d.entry_stride = 2 * ly.ch_off, 2 * self.total_C
                d.cnt_off, d.cnt_stride = i, n_active
        return arr

    @staticmethod
    def _upload(arr):
        return pinned_bytes(arr)

    def finalize(self):
        if self.finalized:
            return
        active = [ly for ly in self.layers if ly.fired]
        if not active:
            self.finalized = True
            return
        dev = self.device
        any_mean_meter = any(not ly.moving_avg for ly in active)
        ids = tuple(ly.idx for ly in active)
        if self.desc_dirty or ids != self._desc_active or any_mean_meter:
            self._desc_host = self._upload(self._build_descs(active))
            self._desc_dev = self._desc_host.to(dev, non_blocking=True)
            if self.process_group is not None:
                ws = torch.distributed.get_world_size(self.process_group)
                self._desc_gath_host = self._upload(self._build_descs(active, (ws, len(active))))
                self._desc_gath = self._desc_gath_host.to(dev, non_blocking=True)
            if ids != self._desc_active:
                self.loss.zero_()       # the ticket slot moves with the number of active layers
            self._desc_active = ids
            self.desc_dirty = False
        st = stream_ptr()
        n = len(active)
        if self.process_group is None:
            call("vitta_stats_finalize", ptr(self._desc_dev), n, ptr(self._base), None, ptr(self.src_mean),
                 ptr(self.src_var), ptr(self.ema_mean), ptr(self.ema_var), ptr(self.batch_mean), ptr(self.batch_var),
                 ptr(self.coef_a), ptr(self.coef_b), ptr(self.loss), 0, None, None, self.max_C, st)
        else:
            pay, merged, cnts = stats_payload(self.total_C, n, dev)
            call("vitta_stats_finalize", ptr(self._desc_dev), n, ptr(self._base), None, None, None, None, None, None,
                 None, None, None, None, 1, ptr(merged), ptr(cnts), self.max_C, st)
            means, counts = gather_stats_payload(pay, self.total_C, self.process_group)
            self._gath_keep = (means, counts)
            call("vitta_stats_finalize", ptr(self._desc_gath), n, ptr(means), ptr(counts), ptr(self.src_mean),
                 ptr(self.src_var), ptr(self.ema_mean), ptr(self.ema_var), ptr(self.batch_mean), ptr(self.batch_var),
                 ptr(self.coef_a), ptr(self.coef_b), ptr(self.loss), 0, None, None, self.max_C, st)
        self._finish(active)

    def _meter_n(self, ly):
        return int(self.global_clips) if (self.process_group is not None and self.global_clips) else ly.n_batch

    def _finish(self, active):
        for ly in active:
            if not ly.moving_avg:
                ly.meter_count += self._meter_n(ly)
        self._loss_index = {ly.idx: i for i, ly in enumerate(active)}
        self._n_active = len(active)
        self.finalized = True
        self.finalize_calls += 1

    def finalize_idle(self, dev):
        """This rank holds no video of the step's global batch (ragged tail with fewer videos than ranks): take part in
        collective C1 with an all-zero payload (count 0 entries are skipped by the Chan merge) and run the same gathered
        finalize as every other rank, so that meters, loss and coefficients stay identical everywhere."""
        if self.process_group is None:
            raise _lib.VittaError("finalize_idle: only meaningful with a process group")
        self._freeze(dev)
        active = list(self.layers)           # every registered layer fires in a forward pass
        n = len(active)
        ws = torch.distributed.get_world_size(self.process_group)
        self._desc_gath_host = self._upload(self._build_descs(active, (ws, n)))
        self._desc_gath = self._desc_gath_host.to(dev, non_blocking=True)
        ids = tuple(ly.idx for ly in active)
        if ids != self._desc_active:
            self.loss.zero_()
            self._desc_active = ids
        self.desc_dirty = True               # the local (non-gathered) descriptors were not built
        pay, merged, cnts = stats_payload(self.total_C, n, dev)
        pay.zero_()
        means, counts = gather_stats_payload(pay, self.total_C, self.process_group)
        self._gath_keep = (means, counts)
        call("vitta_stats_finalize", ptr(self._desc_gath), n, ptr(means), ptr(counts), ptr(self.src_mean),
             ptr(self.src_var), ptr(self.ema_mean), ptr(self.ema_var), ptr(self.batch_mean), ptr(self.batch_var),
             ptr(self.coef_a), ptr(self.coef_b), ptr(self.loss), 0, None, None, self.max_C, stream_ptr())
        self._finish(active)

    def layer_loss(self, ly):
        """r_feature of one layer: differentiable w.r.t. the layer's token (see _RFeature)."""
        self.finalize()
        i = self._loss_index[ly.idx]
        if ly.token is None or not ly.token.requires_grad:
            return self.loss[i].clone()
        return _RFeature.apply(ly.token, self.loss, i)

    def total_loss(self, layers):
        """Sum of r_feature over ``layers`` when they are exactly the layers of this step, as ONE autograd node.

        The finalize kernel already adds the per-layer losses in layer order (fp32, starting from 0) -- the arithmetic
        of the driver's ``loss_reg += hook.r_feature`` loop (reference corpus/basics.py:659-661) -- so the 2 x 47 tiny
        clone / add launches of that loop and their backward nodes collapse into one read.  Returns None when the layer
        set does not match (the caller then sums the hooks one by one)."""
        self.finalize()
        if [ly.idx for ly in layers] != sorted(self._loss_index) or len(layers) != self._n_active:
            return None
        toks = [ly.token for ly in layers]
        if any(t is None or not t.requires_grad for t in toks):
            return None
        return _RTotal.apply(self.loss, self._n_active, *toks)

    def vec(self, t, ly):
        return t[ly.ch_off:ly.ch_off + ly.C]

    def coef_ptrs(self, ly):
        """(coef_a, coef_b, batch_mean) of a layer: dLoss/dy = a + b * (y - mean)."""
        off = ly.ch_off * 4
        return (C.c_void_p(self.coef_a.data_ptr() + off), C.c_void_p(self.coef_b.data_ptr() + off),
                C.c_void_p(self.batch_mean.data_ptr() + off))


class _RFeature(torch.autograd.Function):
    """Ties the scalar loss of a layer (written by the finalize kernel) to that layer's autograd token."""

    @staticmethod
    def forward(ctx, token, loss, i):
        return loss[i].clone()

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class _RTotal(torch.autograd.Function):
    """Total alignment loss (slot n_layers of the finalize kernel's loss vector) tied to every layer's token."""

    @staticmethod
    def forward(ctx, loss, n, *tokens):
        ctx.n_tok = len(tokens)
        return loss[n].clone()

    @staticmethod
    def backward(ctx, g):
        return (None, None) + (g,) * ctx.n_tok


def new_token(ref):
    """0-dim placeholder returned by the forward operators; its gradient is d(total loss)/d(r_feature)."""
    return torch.empty((), dtype=torch.float32, device=ref.device)


# ----------------------------------------------------------------------------------------------
# K1 + K3 on stock modules: statistics tap with closed-form backward
# ----------------------------------------------------------------------------------------------
class StatsTapFn(torch.autograd.Function):
    """feature described as (O, C, I) -> token.  backward: dL/dfeature = g * (a_c + b_c * y).

    ``y`` is ``feature`` itself, or -- for an eval-mode BatchNorm whose output a following in-place ReLU
    overwrites (torchvision Bottleneck) -- recomputed as yscale[c]*saved + yshift[c] from the BatchNorm
    *input* ``saved``.  The gradient is always returned for ``feature`` (the norm output), so autograd's own
    norm backward distributes it to the norm input and affine parameters exactly as in the reference."""

    @staticmethod
    def forward(ctx, feature, saved, arena, ly, O, Cc, I, frames, yscale, yshift):
        arena.record(ly, feature, O, Cc, I, frames)
        ctx.arena, ctx.ly, ctx.geom = arena, ly, (O, Cc, I)
        ctx.save_for_backward(saved, yscale, yshift)
        return new_token(feature)

    @staticmethod
    def backward(ctx, g):
        saved, yscale, yshift = ctx.saved_tensors
        O, Cc, I = ctx.geom
        ca, cb, cm = ctx.arena.coef_ptrs(ctx.ly)
        g = g.contiguous()
        gy = torch.empty_like(saved)
        call("vitta_stats_inject", ptr(saved), ptr(yscale), ptr(yshift), ca, cb, cm, ptr(g), ptr(gy), O, Cc, I,
             stream_ptr())
        return gy, None, None, None, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# K4: fused BN(eval) + stats + residual + ReLU + pooling
# ----------------------------------------------------------------------------------------------
_ws_cache = {}


def _bwd_ws(frames, frame_rows, Cc, dev):
    key = (frames, frame_rows, Cc, dev)
    ws = _ws_cache.get(key)
    if ws is None:
        n = _lib.load().vitta_bn_act_bwd_ws_floats(frames, frame_rows, Cc)
        ws = torch.zeros(n, dtype=torch.float32, device=dev)
        _ws_cache[key] = ws
    return ws


class BNActFn(torch.autograd.Function):
    """out = relu?( BN(x) + [res | BN2(res)] ), statistics partials of BN(x) / BN2(res) into the arena,
    optional per-frame mean of out.  x, res: logical (F, C, H, W), channels_last memory."""

    @staticmethod
    def forward(ctx, x, w, b, rm, rv, eps, res, w2, b2, rm2, rv2, eps2, relu, arena, ly_main, ly_res, want_pool,
                stat_frames):
        _require_cuda(x, "bn_act")
        if not x.is_contiguous(memory_format=CL):
            raise _lib.VittaError("bn_act: x must be channels_last contiguous")
        frames, frame_rows, Cc = as_rows_cl(x)
        dev = x.device
        out = torch.empty_like(x)          # preserves channels_last
        has_res_bn = w2 is not None
        if res is not None and not res.is_contiguous(memory_format=CL):
            raise _lib.VittaError("bn_act: residual must be channels_last contiguous")
        # the kernel's chunking: per-frame when pooling is requested, else one flat frame
        kf, kr = (frames, frame_rows) if want_pool else (1, frames * frame_rows)
        part_main = part_res = None
        if ly_main is not None:
            part_main = arena.partial_buffer(ly_main, frames * frame_rows, Cc, 1, kf, dev)
        if ly_res is not None:
            part_res = arena.partial_buffer(ly_res, frames * frame_rows, Cc, 1, kf, dev)
        pool_part = pool_out = None
        if want_pool:
            ch = _lib.chunking(frames * frame_rows, Cc, 1, kf)
            pool_part = torch.empty(ch.n_entries * Cc, dtype=torch.float32, device=dev)
            pool_out = torch.empty(frames, Cc, dtype=torch.float32, device=dev)
        bn = _lib.make_bn(w, b, rm, rv, eps)
        bn2 = _lib.make_bn(w2, b2, rm2, rv2, eps2) if has_res_bn else None
        if _fused_amax():
            # opt-in f16x3 path: the kernel also emits max|out| for the fp16-split convolution that consumes `out`
            am = new_amax(dev)
            call("vitta_bn_act_fwd_amax", ptr(x), bn, ptr(res), C.byref(bn2) if bn2 is not None else None, int(relu),
                 ptr(out), ptr(part_main), ptr(part_res), ptr(pool_part), ptr(pool_out), kf, kr, Cc, ptr(am), stream_ptr())
            _attach_amax(out, am)
        else:
            call("vitta_bn_act_fwd", ptr(x), bn, ptr(res), C.byref(bn2) if bn2 is not None else None, int(relu),
                 ptr(out), ptr(part_main), ptr(part_res), ptr(pool_part), ptr(pool_out), kf, kr, Cc, stream_ptr())
        ctx.save_for_backward(x, w, b, rm, rv, res, w2, b2, rm2, rv2)
        ctx.meta = (eps, eps2, bool(relu), arena, ly_main, ly_res, want_pool, kf, kr, Cc)
        tok_main = new_token(x) if ly_main is not None else None
        tok_res = new_token(x) if ly_res is not None else None
        return out, pool_out, tok_main, tok_res

    @staticmethod
    def backward(ctx, gout, gpool, gtok_main, gtok_res):
        x, w, b, rm, rv, res, w2, b2, rm2, rv2 = ctx.saved_tensors
        eps, eps2, relu, arena, ly_main, ly_res, want_pool, kf, kr, Cc = ctx.meta
        dev = x.device
        if gout is None:
            gout = torch.zeros_like(x)
        gout = gout.contiguous(memory_format=CL)
        has_res, has_res_bn = res is not None, w2 is not None
        gx = torch.empty_like(x)
        gres = torch.empty_like(res) if has_res else None
        gparam = torch.zeros(4 if has_res_bn else 2, Cc, dtype=torch.float32, device=dev)     # one fill, not four
        gw, gb = gparam[0], gparam[1]
        gw2, gb2 = (gparam[2], gparam[3]) if has_res_bn else (None, None)
        ca = cb = cm = gs = ca2 = cb2 = cm2 = gs2 = None
        if ly_main is not None and gtok_main is not None:
            ca, cb, cm = arena.coef_ptrs(ly_main)
            gs = ptr(gtok_main.contiguous())
        if ly_res is not None and gtok_res is not None:
            ca2, cb2, cm2 = arena.coef_ptrs(ly_res)
            gs2 = ptr(gtok_res.contiguous())
        if gpool is not None:
            gpool = gpool.contiguous()
        ws = _bwd_ws(kf, kr, Cc, dev)
        bn = _lib.make_bn(w, b, rm, rv, eps)
        bn2 = _lib.make_bn(w2, b2, rm2, rv2, eps2) if has_res_bn else None
        if _fused_amax():
            amx = new_amax(dev)
            amr = new_amax(dev) if has_res else None
            call("vitta_bn_act_bwd_amax", ptr(gout), ptr(gpool) if want_pool else None, ptr(x), bn, ptr(res),
                 C.byref(bn2) if bn2 is not None else None, int(relu), ca, cb, cm, gs, ca2, cb2, cm2, gs2, ptr(gx),
                 ptr(gres), ptr(gw), ptr(gb), ptr(gw2), ptr(gb2), ptr(ws), kf, kr, Cc, ptr(amx), ptr(amr), stream_ptr())
            _attach_amax(gx, amx)
            if has_res:
                _attach_amax(gres, amr)
        else:
            call("vitta_bn_act_bwd", ptr(gout), ptr(gpool) if want_pool else None, ptr(x), bn, ptr(res),
                 C.byref(bn2) if bn2 is not None else None, int(relu), ca, cb, cm, gs, ca2, cb2, cm2, gs2, ptr(gx),
                 ptr(gres), ptr(gw), ptr(gb), ptr(gw2), ptr(gb2), ptr(ws), kf, kr, Cc, stream_ptr())
        return (gx, gw, gb, None, None, None, gres, gw2, gb2, None, None, None, None, None, None, None, None, None)


def bn_act(x, bn, relu, res=None, res_bn=None, arena=None, ly_main=None, ly_res=None, want_pool=False):
    """Functional wrapper.  ``bn`` / ``res_bn`` are nn.BatchNorm2d modules in eval mode."""
    w2 = b2 = rm2 = rv2 = None
    eps2 = 0.0
    if res_bn is not None:
        w2, b2, rm2, rv2, eps2 = res_bn.weight, res_bn.bias, res_bn.running_mean, res_bn.running_var, res_bn.eps
    out, pool, tok_main, tok_res = BNActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, res,
                                                 w2, b2, rm2, rv2, eps2, relu, arena, ly_main, ly_res, want_pool, None)
    if ly_main is not None:
        ly_main.token = tok_main
    if ly_res is not None:
        ly_res.token = tok_res
    return out, pool


# ----------------------------------------------------------------------------------------------
# K5: TAM stencil
# ----------------------------------------------------------------------------------------------
_fin_tickets = {}      # per device: the zeroed ticket counters of vitta_tam_bwd_finish (left at zero by every launch)


class TamStencilFn(torch.autograd.Function):
    """x (N*T, C, H, W) channels_last; kern (N, 3, C); act (N, T, C) -> out like x."""

    @staticmethod
    def forward(ctx, x, kern, act, T):
        _require_cuda(x, "tam")
        if not x.is_contiguous(memory_format=CL):
            raise _lib.VittaError("tam: x must be channels_last contiguous")
        nt, Cc, h, w = x.shape
        n = nt // T
        kern, act = kern.contiguous(), act.contiguous()
        out = torch.empty_like(x)
        if _fused_amax():
            am = new_amax(x.device)
            call("vitta_tam_fwd_amax", ptr(x), ptr(kern), ptr(act), ptr(out), n, T, h * w, Cc, ptr(am), stream_ptr())
            _attach_amax(out, am)
        else:
            call("vitta_tam_fwd", ptr(x), ptr(kern), ptr(act), ptr(out), n, T, h * w, Cc, stream_ptr())
        ctx.save_for_backward(x, kern, act)
        ctx.T = T
        return out

    @staticmethod
    def backward(ctx, gout):
        x, kern, act = ctx.saved_tensors
        T = ctx.T
        nt, Cc, h, w = x.shape
        n = nt // T
        gout = gout.contiguous(memory_format=CL)
        gx = torch.empty_like(x)
        nch = _lib.load().vitta_tam_num_chunks(h * w, Cc)
        dpart = torch.empty(n, nch, T, 3, Cc, dtype=torch.float32, device=x.device)
        call("vitta_tam_bwd", ptr(gout), ptr(x), ptr(kern), ptr(act), ptr(gx), ptr(dpart), n, T, h * w, Cc, stream_ptr())
        if 3 <= T <= 16:
            gkern = torch.empty_like(kern)                 # (N, 3, C)
            gact = torch.empty_like(act)                   # (N, T, C)
            tk = _fin_tickets.get(x.device)
            if tk is None:
                tk = _fin_tickets[x.device] = torch.zeros(_lib.load().vitta_tam_bwd_finish_tickets(), dtype=torch.int32,
                                                          device=x.device)
            call("vitta_tam_bwd_finish", ptr(dpart), ptr(kern), ptr(act), ptr(gkern), ptr(gact), ptr(tk), n, T, nch, Cc,
                 stream_ptr())
        else:
            D = dpart.sum(1)                               # (N, T, 3, C): tiny
            gkern = (act.unsqueeze(2) * D).sum(1)
            gact = (kern.unsqueeze(1) * D).sum(2)
        return gx, gkern, gact, None


# ----------------------------------------------------------------------------------------------
# K10: prediction consistency
# ----------------------------------------------------------------------------------------------
class PredConsisFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds):
        _require_cuda(preds, "pred_consis")
        p = preds.contiguous()
        b, v, k = p.shape
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        call("vitta_pred_consis", ptr(p), b, v, k, ptr(loss), ptr(grad), stream_ptr())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        grad, = ctx.saved_tensors
        return grad * g


# ----------------------------------------------------------------------------------------------
# K11: fused SGD
# ----------------------------------------------------------------------------------------------
class FusedSGD:
    """torch.optim.SGD(params, lr, momentum, weight_decay) semantics (dampening 0, no nesterov) in one
    kernel launch over a device-resident tensor table (corpus/basics.py:559-560,669-671).  Parameters
    whose ``.grad`` is None are skipped exactly like torch does (no decay, no momentum update)."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, process_group=None):
        self.params = [p for p in params]
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.bufs = {}
        self._key = None
        self._tables = None
        self.process_group = process_group
        self.param_groups = [{"params": self.params, "lr": self.lr, "momentum": self.momentum,
                              "weight_decay": self.weight_decay}]
        self._block = None
        self._pin = []
        self._pin_next = 0
        self._bucketer = None
        self._bucket_live = None
        self.overlap_allreduce = os.environ.get("VITTA_OVERLAP_ALLREDUCE", "1") == "1"
        self._pin_cap = 0

    def set_overlap(self, on):
        """Switch the bucketed all-reduce of this step on / off on EVERY rank alike (off for ragged steps)."""
        if self._bucketer is not None:
            self._bucketer.enabled = bool(on)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def _table(self, items, dev):
        n = len(items)
        arr = (_lib.VittaSgdTensor * n)()
        starts = []
        blk = 0
        for i, (p, g, buf) in enumerate(items):
            arr[i].p, arr[i].g, arr[i].buf, arr[i].n = p.data_ptr(), g.data_ptr(), buf.data_ptr(), p.numel()
            starts.append(blk)
            blk += (p.numel() + self._block - 1) // self._block
        # Pinned staging buffers allocated ONCE (no host allocation may happen while a CUDA graph is being captured) and
        # non-blocking copies (memcpy nodes inside a capture).  Eager steps rotate through slots 0-3 so that the previous
        # tables' sources stay intact; tables built DURING a capture use the reserved slots 4-5, which no eager step
        # ever rewrites -- a graph replay re-reads its memcpy node's pinned source every time (after four
        # eager steps a captured slot of the rotation would have held stale gradient pointers).
        nbytes = C.sizeof(_lib.VittaSgdTensor) * len(self.params)
        if not self._pin:
            self._pin = [(torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                          torch.empty(len(self.params), dtype=torch.int32).pin_memory()) for _ in range(6)]
        if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
            tab_h, st_h = self._pin[4 + self._pin_cap % 2]
            self._pin_cap += 1
        else:
            tab_h, st_h = self._pin[self._pin_next % 4]
            self._pin_next += 1
        raw = bytes(arr)
        C.memmove(tab_h.data_ptr(), raw, len(raw))
        st_np = (C.c_int32 * n)(*starts)
        C.memmove(st_h.data_ptr(), st_np, 4 * n)
        return (tab_h[:len(raw)].to(dev, non_blocking=True), st_h[:n].to(dev, non_blocking=True), n, blk)

    def exchange_live_mask(self):
        """Ragged global batch (some rank holds no video): group rank 0 -- which always holds one -- tells every rank
        which parameters received a gradient, so that idle ranks can contribute zeros of the right layout to C2."""
        import torch.distributed as dist
        dev = self.params[0].device
        mask = torch.tensor([1 if p.grad is not None else 0 for p in self.params], dtype=torch.uint8, device=dev)
        dist.broadcast(mask, src=dist.get_global_rank(self.process_group, 0), group=self.process_group)
        return [bool(v) for v in mask.tolist()]

    @torch.no_grad()
    def step_idle(self):
        """The step of a rank without videos: zero gradients for rank 0's live set, then the common all-reduce + update
        (the weights must move exactly as on the ranks that did the work)."""
        live = self.exchange_live_mask()
        for p, on in zip(self.params, live):
            p.grad = torch.zeros_like(p) if on else None
        self.step()
        self.zero_grad()

    @torch.no_grad()
    def step(self):
        if self._block is None:
            self._block = _lib.load().vitta_sgd_block_elems()
        lr = float(self.param_groups[0]["lr"])
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        dev = live[0].device
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in live]
        if self.process_group is not None:
            views = self._bucketer.finish(live) if self._bucketer is not None else None
            if views is not None:
                grads = [views[p] for p in live]          # collective C2 ran in buckets DURING the backward pass
            else:
                grads, self._flat_keep = allreduce_grads(grads, self.process_group)   # collective C2, one flat buffer
                if self.overlap_allreduce and self._bucket_live != [id(p) for p in live]:
                    # (re)build the buckets for the live set just seen; they take over from the next step on
                    if self._bucketer is not None:
                        self._bucketer.close()
                    self._bucketer = GradBucketer(live, self.process_group)
                    self._bucket_live = [id(p) for p in live]
        first, rest = [], []
        for p, g in zip(live, grads):
            if not p.is_contiguous():
                raise _lib.VittaError("FusedSGD: parameters must be contiguous")
            buf = self.bufs.get(p)
            if buf is None:
                buf = torch.empty_like(p, memory_format=torch.contiguous_format)
                self.bufs[p] = buf
                first.append((p, g, buf))
            else:
                rest.append((p, g, buf))
        key = (tuple((p.data_ptr(), g.data_ptr()) for p, g, _ in first), tuple((p.data_ptr(), g.data_ptr()) for p, g, _ in rest))
        if key != self._key:
            self._tables = (self._table(first, dev) if first else None, self._table(rest, dev) if rest else None)
            self._key = key
        st = stream_ptr()
        for tab, is_first in ((self._tables[0], 1), (self._tables[1], 0)):
            if tab is None:
                continue
            t, starts, n, blocks = tab
            call("vitta_sgd_step", ptr(t), ptr(starts), n, blocks, lr, self.momentum, self.weight_decay, is_first, 1.0, st)
        bump_weight_epoch()
        refresh_weight_splits(self.params)       # all updated weights, both operand forms, one launch sequence


# ----------------------------------------------------------------------------------------------
# K6/K8: tcgen05 3xTF32 GEMM / implicit-GEMM convolution
# ----------------------------------------------------------------------------------------------
def split_tf32(w, mode=0):
    """Weight preparation (vitta_split_tf32).  ``w``: Linear weight (N, K) or conv weight (Cout, Cin, KH, KW) in
    channels_last memory.  mode 0 -> forward operand [Cout][tap][Cin]; mode 1 -> data-gradient operand
    [Cin][rotated tap][Cout].  Returns (hi, lo) flat fp32 tensors."""
    _require_cuda(w, "split_tf32")
    if w.dim() == 2:
        r, t, c = w.shape[0], 1, w.shape[1]
        src = w.contiguous()
    else:
        r, c, kh, kw = w.shape
        t = kh * kw
        src = w.contiguous(memory_format=CL) if t > 1 else w.reshape(r, c).contiguous()
    hi = torch.empty(r * t * c, dtype=torch.float32, device=w.device)
    lo = torch.empty_like(hi)
    call("vitta_split_tf32", ptr(src), ptr(hi), ptr(lo), r, t, c, int(mode), stream_ptr())
    return hi, lo


def gemm_tf32x3(a, b_hi, b_lo, n, bias=None, residual=None, act=0, out=None, force_bn=0):
    """out[M, n] = a[M, K] @ B[n, K]^T (+bias) (GELU if act=1) (+residual); B given pre-split (split_tf32)."""
    _require_cuda(a, "gemm_tf32x3")
    if a.dim() != 2 or a.stride(1) != 1:
        raise _lib.VittaError("gemm_tf32x3: A must be (M, K) with unit inner stride")
    m, k = a.shape
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    ldr = residual.stride(0) if residual is not None else 0
    call("vitta_gemm_tf32x3", ptr(a), a.stride(0), ptr(b_hi), ptr(b_lo), k, ptr(out), out.stride(0), m, n, k,
         ptr(bias), ptr(residual), ldr, int(act), int(force_bn), stream_ptr())
    return out


def conv2d_tf32x3(x, w_hi, w_lo, cout, kh, kw, stride, pad, bias=None, force_bn=0, residual=None):
    """x: logical (F, Cin, H, W) in channels_last memory -> logical (F, cout, Ho, Wo) channels_last.
    ``residual`` (shape of the result, channels_last) is added in the epilogue."""
    _require_cuda(x, "conv2d_tf32x3")
    if not x.is_contiguous(memory_format=CL):
        raise _lib.VittaError("conv2d_tf32x3: x must be channels_last contiguous")
    f, cin, h, w = x.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    y = torch.empty((f, cout, ho, wo), dtype=torch.float32, device=x.device, memory_format=CL)
    if residual is not None and not (residual.shape == y.shape and residual.is_contiguous(memory_format=CL)):
        raise _lib.VittaError("conv2d_tf32x3: residual must be channels_last with the shape of the result")
    call("vitta_conv2d_tf32x3_ex", ptr(x), f, h, w, cin, ptr(w_hi), ptr(w_lo), cout, kh, kw, stride, pad, ptr(y), ptr(bias),
         ptr(residual), int(force_bn), stream_ptr())
    return y


# ---- fp16 operand split (kind::f16): experimental in round 1 -- exported, opt-in tested, not yet used by the step ----
def amax_f32(x, out=None):
    """Device scalar max|x| (vitta_amax_f32).  ``out``: an existing scalar to accumulate into (max), else a fresh zero."""
    _require_cuda(x, "amax_f32")
    if not x.is_contiguous() and not x.is_contiguous(memory_format=CL):
        raise _lib.VittaError("amax_f32: x must be dense")
    if out is None:
        out = new_amax(x.device)
    call("vitta_amax_f32", ptr(x), x.numel(), ptr(out), stream_ptr())
    return out


def split_f16(w, mode=0):
    """fp16 weight preparation (vitta_split_f16): same shapes / modes as split_tf32.  Returns (hi, lo, amax): two flat
    fp16 tensors holding w * s and the device scalar the scale s is derived from."""
    _require_cuda(w, "split_f16")
    if w.dim() == 2:
        r, t, c = w.shape[0], 1, w.shape[1]
        src = w.contiguous()
    else:
        r, c, kh, kw = w.shape
        t = kh * kw
        src = w.contiguous(memory_format=CL) if t > 1 else w.reshape(r, c).contiguous()
    # a weight's range scalar lives as long as the cached split: never a slot of the per-pass pool
    am = amax_f32(src, out=torch.zeros(1, dtype=torch.float32, device=w.device))
    hi = torch.empty(r * t * c, dtype=torch.float16, device=w.device)
    lo = torch.empty_like(hi)
    call("vitta_split_f16", ptr(src), ptr(hi), ptr(lo), ptr(am), r, t, c, int(mode), stream_ptr())
    return hi, lo, am


def gemm_f16x3(a, b_hi, b_lo, b_amax, n, a_amax=None, bias=None, residual=None, act=0, out=None, force_bn=0):
    """gemm_tf32x3 on the fp16 split; ``a_amax`` defaults to a fresh vitta_amax_f32 pass over ``a``."""
    _require_cuda(a, "gemm_f16x3")
    if a.dim() != 2 or a.stride(1) != 1:
        raise _lib.VittaError("gemm_f16x3: A must be (M, K) with unit inner stride")
    m, k = a.shape
    if a_amax is None:
        a_amax = amax_f32(a if a.is_contiguous() else a.contiguous())
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    ldr = residual.stride(0) if residual is not None else 0
    call("vitta_gemm_f16x3_ex", ptr(a), a.stride(0), ptr(a_amax), ptr(b_hi), ptr(b_lo), ptr(b_amax), k, ptr(out),
         out.stride(0), m, n, k, ptr(bias), ptr(residual), ldr, int(act), None, None, 1, int(force_bn), stream_ptr())
    return out


def conv2d_f16x3(x, w_hi, w_lo, w_amax, cout, kh, kw, stride, pad, x_amax=None, bias=None, force_bn=0, residual=None):
    """conv2d_tf32x3 on the fp16 split (channels_last in, channels_last out)."""
    _require_cuda(x, "conv2d_f16x3")
    if not x.is_contiguous(memory_format=CL):
        raise _lib.VittaError("conv2d_f16x3: x must be channels_last contiguous")
    f, cin, h, w = x.shape
    if x_amax is None:
        x_amax = amax_f32(x)
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    y = torch.empty((f, cout, ho, wo), dtype=torch.float32, device=x.device, memory_format=CL)
    call("vitta_conv2d_f16x3_ex", ptr(x), ptr(x_amax), f, h, w, cin, ptr(w_hi), ptr(w_lo), ptr(w_amax), cout, kh, kw,
         stride, pad, ptr(y), ptr(bias), ptr(residual), int(force_bn), stream_ptr())
    return y


def conv2d_dgrad_f16x3(gy, wt_hi, wt_lo, w_amax, x_shape, kh, kw, stride, pad, gy_amax=None):
    """Data gradient of a (strided) convolution on the fp16 split; wt_* from split_f16(w, mode=1)."""
    _require_cuda(gy, "conv2d_dgrad_f16x3")
    f, cin, h, w = x_shape
    cout = gy.shape[1]
    if gy_amax is None:
        gy_amax = amax_f32(gy)
    gx = torch.empty(x_shape, dtype=torch.float32, device=gy.device).contiguous(memory_format=CL)
    call("vitta_conv2d_dgrad_f16x3", ptr(gy), ptr(gy_amax), f, gy.shape[2], gy.shape[3], cout, ptr(wt_hi), ptr(wt_lo),
         ptr(w_amax), cin, kh, kw, stride, pad, h, w, ptr(gx), stream_ptr())
    return gx


# weight-split cache: the hi/lo operands of a weight are rebuilt only when the weight changed.  The entry lives ON the
# weight tensor object (like _vitta_amax above), so it dies with the tensor: a freed weight whose address the caching
# allocator hands to a new same-shape weight can never serve its stale split (round-1 bug: the cache was keyed by
# data_ptr).  autograd returns saved INPUT tensors as the original objects, so backward() finds the forward's entry; a
# re-wrapped tensor merely misses and splits again.  Validated by the autograd version counter (in-place updates,
# load_state_dict) and by a global epoch that every FusedSGD step / graph replay bumps (they update parameters through
# raw pointers, invisible to the version counter).
_weight_epoch = 0
_split_registry = {}     # id(weight) -> weakref(weight): weights that carry cached splits (for the multi-tensor refresh)


def bump_weight_epoch():
    global _weight_epoch
    _weight_epoch += 1


def _stamp(w):
    return (w._version, _weight_epoch, w.data_ptr(), tuple(w.shape))


def _cached_split(w, mode, kind, make):
    cache = getattr(w, "_vitta_split", None)
    if cache is None:
        cache = {}
        try:
            w._vitta_split = cache
        except AttributeError:       # an object without a __dict__: no caching, always correct
            cache = None
    stamp = _stamp(w)
    ent = cache.get((mode, kind)) if cache is not None else None
    if ent is None or ent[0] != stamp:
        with torch.no_grad():
            ent = [stamp, make(w.detach(), mode)]
        if cache is not None:
            cache[(mode, kind)] = ent
            if id(w) not in _split_registry:
                import weakref
                key = id(w)
                _split_registry[key] = weakref.ref(w, lambda _r, k=key: _split_registry.pop(k, None))
    return ent[1]


class _SplitTable:
    """Device table of vitta_split_multi for one set of (weight, operand form) entries; rebuilt only when the set or any
    buffer address changes (never inside a steady-state step, so it is CUDA-graph safe)."""

    def __init__(self):
        self.key = None
        self.dev = None
        self._keep = []      # superseded tables stay alive: a captured CUDA graph may still read them

    def build(self, items, dev):
        if self.dev is not None:
            self._keep.append((self.dev, self._host))
        blk = _lib.load().vitta_split_block_elems()
        arr = (_lib.VittaSplitTensor * len(items))()
        starts, b = [], 0
        for i, (w, mode, bufs, geom) in enumerate(items):
            r, t, c, tap_inner = geom
            e = arr[i]
            e.src, e.hi, e.lo = w.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr()
            e.amax = bufs[2].data_ptr() if len(bufs) > 2 else None
            e.R, e.T, e.Cc, e.mode, e.src_tap_inner = r, t, c, mode, tap_inner
            e.compute_amax = 1 if len(bufs) > 2 else 0
            e.n = r * t * c
            starts.append(b)
            b += (e.n + blk - 1) // blk
        self._host = (pinned_bytes(arr), torch.tensor(starts, dtype=torch.int32).pin_memory())
        self.dev = (self._host[0].to(dev, non_blocking=True), self._host[1].to(dev, non_blocking=True), len(items), b)


_split_tables = {"f16": _SplitTable(), "tf32": _SplitTable()}


def _split_geom(w):
    """(R, T, Cc, src_tap_inner) of a weight the multi-tensor kernel can read in place, else None."""
    if not w.is_contiguous():
        return None
    if w.dim() == 2:
        return int(w.shape[0]), 1, int(w.shape[1]), 0
    if w.dim() == 4:
        r, c, kh, kw = (int(v) for v in w.shape)
        return r, kh * kw, c, (1 if kh * kw > 1 else 0)
    return None


def refresh_weight_splits(params=None, launch=True):
    """Bring the cached operand splits of all registered weights (restricted to ``params`` when given) up to date with
    ONE multi-tensor launch sequence, writing into the existing buffers, and stamp them valid.  Called by FusedSGD.step
    right after the update (so the next forward / backward finds every split ready), and with ``launch=False`` after a
    CUDA-graph replay (the captured step already ran the refresh; only the stamps have to follow)."""
    kind = "f16" if _gemm_precision == "f16x3" else "tf32"
    allowed = None if params is None else {id(p) for p in params}
    items, dev = [], None
    for key, ref in list(_split_registry.items()):
        w = ref()
        if w is None or (allowed is not None and key not in allowed):
            continue
        geom = _split_geom(w)
        cache = getattr(w, "_vitta_split", None)
        if geom is None or not cache or not w.is_cuda:
            continue
        for (mode, k2), ent in cache.items():
            if k2 == kind and mode in (0, 1):        # other forms (the stem's packed operand) re-split lazily
                items.append((w, mode, ent[1], geom, ent))
                dev = w.device
    if not items:
        return 0
    if launch:
        tab = _split_tables[kind]
        key = tuple((id(w), mode, w.data_ptr()) + tuple(b.data_ptr() for b in bufs) for w, mode, bufs, _, _ in items)
        if tab.key != key:
            tab.build([it[:4] for it in items], dev)
            tab.key = key
        t, starts, n, blocks = tab.dev
        call("vitta_split_multi", ptr(t), ptr(starts), n, blocks, 1 if kind == "f16" else 0, stream_ptr())
    for w, mode, bufs, geom, ent in items:
        ent[0] = _stamp(w)
    return len(items)


# Operand split of the dense contractions: "f16x3" (default since round 2: fp16 hi/lo pieces on
# kind::f16 with per-tensor power-of-two scales from amax -- the same ~2^-21 per-product error as the tf32 split at half
# the tensor-pipe work and half the weight bytes; validated on hardware against float64 and the reference goldens) or
# "tf32x3" (hi/lo tf32 pieces on kind::tf32).  Covers forward, data-gradient and weight-gradient convolutions / Linear
# layers; the window-attention kernels keep the tf32 split.  Settable with VITTA_GEMM_PRECISION / set_gemm_precision().
_gemm_precision = os.environ.get("VITTA_GEMM_PRECISION", "f16x3")
_FUSED_AMAX = os.environ.get("VITTA_FUSED_AMAX", "1") == "1"    # f16x3 only: 0 = always use standalone amax passes


def set_gemm_precision(name):
    global _gemm_precision
    if name not in ("tf32x3", "f16x3"):
        raise _lib.VittaError("gemm precision must be 'tf32x3' or 'f16x3', got %r" % (name,))
    _gemm_precision = name
    bump_weight_epoch()


def gemm_precision():
    return _gemm_precision


def weight_split_f16(w, mode):
    """Cached fp16 pieces + amax scalar of a weight (same invalidation rules as weight_split)."""
    return _cached_split(w, mode, "f16", split_f16)


def weight_split(w, mode):
    """Cached tf32 (hi, lo) operands of a weight; see the cache rules above."""
    return _cached_split(w, mode, "tf32", split_tf32)


_wgrad_ws = {}


def conv2d_wgrad_tf32x3(x, gy, cout, kh, kw, stride, pad):
    """Weight gradient (Cout, Cin, KH, KW), contiguous, of conv2d(x, w, stride, pad) given dL/dy = gy.
    x, gy: channels_last."""
    _require_cuda(x, "conv2d_wgrad")
    if not (x.is_contiguous(memory_format=CL) and gy.is_contiguous(memory_format=CL)):
        raise _lib.VittaError("conv2d_wgrad: x and gy must be channels_last contiguous")
    f, cin, h, w = x.shape
    n = _lib.load().vitta_conv2d_wgrad_ws_floats(f, h, w, cin, cout, kh, kw, stride, pad)
    if n <= 0:
        raise _lib.VittaError("conv2d_wgrad: bad geometry")
    ws = _wgrad_ws.get(x.device)
    if ws is None or ws.numel() < n:
        ws = torch.empty(n, dtype=torch.float32, device=x.device)
        _wgrad_ws[x.device] = ws
    gw = torch.empty((cout, cin, kh, kw), dtype=torch.float32, device=x.device)
    call("vitta_conv2d_wgrad_tf32x3", ptr(x), ptr(gy), f, h, w, cin, cout, kh, kw, stride, pad, ptr(gw), 0, ptr(ws),
         stream_ptr())
    return gw


def conv2d_wgrad_f16x3(x, gy, cout, kh, kw, stride, pad, x_amax=None, gy_amax=None):
    """conv2d_wgrad_tf32x3 on the fp16 split (opt-in)."""
    _require_cuda(x, "conv2d_wgrad_f16x3")
    if not (x.is_contiguous(memory_format=CL) and gy.is_contiguous(memory_format=CL)):
        raise _lib.VittaError("conv2d_wgrad_f16x3: x and gy must be channels_last contiguous")
    f, cin, h, w = x.shape
    n = _lib.load().vitta_conv2d_wgrad_ws_floats(f, h, w, cin, cout, kh, kw, stride, pad)
    if n <= 0:
        raise _lib.VittaError("conv2d_wgrad_f16x3: bad geometry")
    ws = _wgrad_ws.get(x.device)
    if ws is None or ws.numel() < n:
        ws = torch.empty(n, dtype=torch.float32, device=x.device)
        _wgrad_ws[x.device] = ws
    if x_amax is None:
        x_amax = amax_f32(x)
    if gy_amax is None:
        gy_amax = amax_f32(gy)
    gw = torch.empty((cout, cin, kh, kw), dtype=torch.float32, device=x.device)
    call("vitta_conv2d_wgrad_f16x3", ptr(x), ptr(x_amax), ptr(gy), ptr(gy_amax), f, h, w, cin, cout, kh, kw, stride, pad,
         ptr(gw), 0, ptr(ws), stream_ptr())
    return gw


class Conv2dFn(torch.autograd.Function):
    """Bias-free 2-D convolution, channels-last, forward and data gradient on the tcgen05 3xTF32 kernel.

    Stride-1 data gradients reuse the forward kernel with the transposed / 180-degree-rotated operand (split mode 1);
    strided data gradients run the same kernel once per residue class of the input pixels (vitta_conv2d_dgrad_tf32x3);
    weight gradients run on the split-K tcgen05 wgrad kernel."""

    @staticmethod
    def forward(ctx, x, w, stride, pad, want_alias=False):
        cout, cin, kh, kw = w.shape
        ctx.x_am = None
        if _gemm_precision == "f16x3" and (cin * kh * kw) % 8 == 0:
            whi, wlo, wam = weight_split_f16(w, 0)
            ctx.x_am = operand_amax(x)      # reused by the weight gradient (same tensor)
            y = conv2d_f16x3(x, whi, wlo, wam, cout, kh, kw, stride, pad, x_amax=ctx.x_am)
        else:
            whi, wlo = weight_split(w, 0)
            y = conv2d_tf32x3(x, whi, wlo, cout, kh, kw, stride, pad)
        ctx.save_for_backward(x, w)
        ctx.geom = (stride, pad)
        if want_alias:
            # the input also feeds the block's shortcut: hand out an alias whose gradient is added inside the
            # data-gradient kernel's epilogue instead of by a separate autograd accumulation pass
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, gy, galias=None):
        x, w = ctx.saved_tensors
        stride, pad = ctx.geom
        if galias is not None:
            galias = galias.contiguous(memory_format=CL)
        cout, cin, kh, kw = w.shape
        gy = gy.contiguous(memory_format=CL)
        gx = gw = None
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        f16 = _gemm_precision == "f16x3" and (cout * kh * kw) % 8 == 0
        gy_am = operand_amax(gy) if _gemm_precision == "f16x3" else None    # shared by dgrad and wgrad
        if need_x and stride == 1 and kh == kw and x.shape[2:] == gy.shape[2:]:
            if f16:
                whi, wlo, wam = weight_split_f16(w, 1)
                gx = conv2d_f16x3(gy, whi, wlo, wam, cin, kh, kw, 1, kh - 1 - pad, x_amax=gy_am, residual=galias)
            else:
                whi, wlo = weight_split(w, 1)
                gx = conv2d_tf32x3(gy, whi, wlo, cin, kh, kw, 1, kh - 1 - pad, residual=galias)
            galias = None
            need_x = False
        elif need_x and stride > 1 and kh * kw <= 9 and cout % 4 == 0:
            f, _, h, wd = x.shape
            gx = torch.empty_like(x)             # channels_last like x
            if f16:
                whi, wlo, wam = weight_split_f16(w, 1)
                call("vitta_conv2d_dgrad_f16x3", ptr(gy), ptr(gy_am), f, gy.shape[2], gy.shape[3], cout, ptr(whi),
                     ptr(wlo), ptr(wam), cin, kh, kw, stride, pad, h, wd, ptr(gx), stream_ptr())
            else:
                whi, wlo = weight_split(w, 1)
                call("vitta_conv2d_dgrad_tf32x3", ptr(gy), f, gy.shape[2], gy.shape[3], cout, ptr(whi), ptr(wlo), cin, kh,
                     kw, stride, pad, h, wd, ptr(gx), stream_ptr())
            need_x = False
        if need_w and cout % 4 == 0:
            if gy_am is not None:
                gw = conv2d_wgrad_f16x3(x, gy, cout, kh, kw, stride, pad, x_amax=ctx.x_am, gy_amax=gy_am)
            else:
                gw = conv2d_wgrad_tf32x3(x, gy, cout, kh, kw, stride, pad)
            need_w = False
        if need_x or need_w:
            r = torch.ops.aten.convolution_backward(gy, x, w, None, [stride, stride], [pad, pad], [1, 1], False, [0, 0],
                                                    1, [need_x, need_w, False])
            if need_x:
                gx = r[0]
            if need_w:
                gw = r[1]
        if galias is not None:
            gx = galias if gx is None else gx + galias
        return gx, gw, None, None, None


def conv2d_shortcut(x, w, stride, pad):
    """conv2d(x, w) and an alias of x for the residual path (see Conv2dFn.forward)."""
    y, alias = Conv2dFn.apply(x, w, stride, pad, True)
    # the alias is a new tensor object over the same values: hand the operand range on, or the block's downsample
    # convolution pays a standalone range pass over a tensor whose range is already known
    ent = getattr(x, "_vitta_amax", None)
    if ent is not None and ent[1] == x._version and ent[2] == _amax_gen:
        _attach_amax(alias, ent[0])
    return y, alias


def conv2d(x, w, stride, pad):
    return Conv2dFn.apply(x, w, stride, pad, False)


# ----------------------------------------------------------------------------------------------
# ResNet stem: conv1 7x7/2 on tcgen05 (overlapping-row TMA operand), BN + ReLU + MaxPool in one pass
# ----------------------------------------------------------------------------------------------
def _stem_weight_split(w, mode):
    hi = torch.empty(64 * 224, dtype=torch.float32, device=w.device)
    lo = torch.empty_like(hi)
    call("vitta_stem_pack_weight", ptr(w.contiguous()), ptr(hi), ptr(lo), stream_ptr())
    return hi, lo


_stem_ws = {}


class StemConvFn(torch.autograd.Function):
    """conv1 of the ResNet trunk (64 x 3 x 7 x 7, stride 2, padding 3) on the tcgen05 3xTF32 kernel.  x: (F, 3, H, W) in
    ANY dense layout (read once by the packing kernel); result (F, 64, H/2, W/2) channels_last.  The weight gradient --
    the only gradient the step needs here, the frames do not require one -- is the register-blocked fp32 kernel
    vitta_stem_wgrad over the same packed image (the library convolution backward is used only for an input gradient)."""

    @staticmethod
    def forward(ctx, x, w):
        _require_cuda(x, "stem_conv")
        f, c, h, wd = x.shape
        xc = x if x.is_contiguous() else x.contiguous()
        xp = torch.empty(f, h + 6, wd + 6, 4, dtype=torch.float32, device=x.device)
        call("vitta_stem_pack", ptr(xc), ptr(xp), f, h, wd, stream_ptr())
        whi, wlo = _cached_split(w, "stem", "tf32", _stem_weight_split)
        y = torch.empty((f, 64, h // 2, wd // 2), dtype=torch.float32, device=x.device, memory_format=CL)
        call("vitta_stem_conv_tf32x3", ptr(xp), f, h, wd, ptr(whi), ptr(wlo), ptr(y), stream_ptr())
        ctx.save_for_backward(xc, w, xp)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, xp = ctx.saved_tensors
        f, _, h, wd = x.shape
        gx = gw = None
        need_x, need_w = ctx.needs_input_grad
        gy = gy.contiguous(memory_format=CL)
        if need_w:
            ws = _stem_ws.get(x.device)
            if ws is None:
                ws = _stem_ws[x.device] = torch.empty(_lib.load().vitta_stem_wgrad_ws_floats(), dtype=torch.float32,
                                                      device=x.device)
            gw = torch.empty(64, 3, 7, 7, dtype=torch.float32, device=x.device)
            call("vitta_stem_wgrad", ptr(xp), ptr(gy), ptr(gw), ptr(ws), f, h, wd, stream_ptr())
        if need_x:
            gx = torch.ops.aten.convolution_backward(gy, x.contiguous(memory_format=CL), w, None, [2, 2], [3, 3], [1, 1],
                                                     False, [0, 0], 1, [True, False, False])[0]
        return gx, gw


def stem_conv_supported(conv, x):
    return (tuple(conv.weight.shape) == (64, 3, 7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None and conv.padding_mode == 'zeros'
            and x.dim() == 4 and x.shape[1] == 3 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and x.shape[2] >= 8
            and x.shape[3] >= 8 and not conv._forward_hooks and not conv._forward_pre_hooks)


_pool_ws = {}


class BnReluPoolFn(torch.autograd.Function):
    """maxpool3x3/2/1(relu(BN_eval(x))) in one pass (reference: torchvision bn1 -> relu -> maxpool inside
    models/tanet_models/tanet.py's base_model); x (F, C, H, W) channels_last.  Saves x and one byte per pooled element
    (the winner's window position); the backward is one pass as well."""

    @staticmethod
    def forward(ctx, x, w, b, rm, rv, eps):
        _require_cuda(x, "bn_relu_pool")
        if not x.is_contiguous(memory_format=CL):
            raise _lib.VittaError("bn_relu_pool: x must be channels_last contiguous")
        f, c, h, wd = x.shape
        ho, wo = (h + 1) // 2, (wd + 1) // 2
        out = torch.empty((f, c, ho, wo), dtype=torch.float32, device=x.device, memory_format=CL)
        code = torch.empty(f * ho * wo * c, dtype=torch.uint8, device=x.device)
        if _fused_amax():
            am = new_amax(x.device)
            call("vitta_bn_relu_pool_fwd_amax", ptr(x), _lib.make_bn(w, b, rm, rv, eps), ptr(out), ptr(code), f, h, wd, c,
                 ptr(am), stream_ptr())
            _attach_amax(out, am)
        else:
            call("vitta_bn_relu_pool_fwd", ptr(x), _lib.make_bn(w, b, rm, rv, eps), ptr(out), ptr(code), f, h, wd, c,
                 stream_ptr())
        ctx.save_for_backward(x, w, b, rm, rv, code)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, gout):
        x, w, b, rm, rv, code = ctx.saved_tensors
        f, c, h, wd = x.shape
        gout = gout.contiguous(memory_format=CL)
        key = (x.device, c)
        ws = _pool_ws.get(key)
        if ws is None:
            ws = _pool_ws[key] = torch.zeros(_lib.load().vitta_bn_relu_pool_bwd_ws_floats(c), dtype=torch.float32,
                                             device=x.device)
        gx = torch.empty_like(x)
        gparam = torch.empty(2, c, dtype=torch.float32, device=x.device)
        call("vitta_bn_relu_pool_bwd", ptr(gout), ptr(code), ptr(x), _lib.make_bn(w, b, rm, rv, ctx.eps), ptr(gx),
             ptr(gparam[0]), ptr(gparam[1]), ptr(ws), f, h, wd, c, stream_ptr())
        return gx, gparam[0], gparam[1], None, None, None


# ----------------------------------------------------------------------------------------------
# K5b: the TAM's G and L gate networks (eval-mode BatchNorm1d) in 2 launches forward / 2 backward
# ----------------------------------------------------------------------------------------------
_gate_ws = {}


class TamGateFn(torch.autograd.Function):
    """pooled (N*T, C) -> (kern (N, 3, C), act (N, T, C)): see include/vitta_b200.h K5b.  Parameters are passed as tensors
    so autograd hands their gradients to the usual accumulation."""

    @staticmethod
    def forward(ctx, pooled, w1, g_w, g_b, g_rm, g_rv, w2, wa, l_w, l_b, l_rm, l_rv, wb, t, eps1, eps2):
        _require_cuda(pooled, "tam_gate")
        pooled = pooled.contiguous()
        nt, c = pooled.shape
        n = nt // t
        dev = pooled.device
        kern = torch.empty(n, 3, c, dtype=torch.float32, device=dev)
        act = torch.empty(n, t, c, dtype=torch.float32, device=dev)
        pre = torch.empty(nt, c // 4, dtype=torch.float32, device=dev)
        w1c, w2c, wac, wbc = w1.contiguous(), w2.contiguous(), wa.contiguous(), wb.contiguous()
        call("vitta_tam_gate_fwd", ptr(pooled), ptr(w1c), _lib.make_bn(g_w, g_b, g_rm, g_rv, eps1), ptr(w2c), ptr(wac),
             _lib.make_bn(l_w, l_b, l_rm, l_rv, eps2), ptr(wbc), ptr(kern), ptr(act), ptr(pre), n, t, c, stream_ptr())
        ctx.save_for_backward(pooled, w1c, g_w, g_b, g_rm, g_rv, w2c, wac, l_w, l_b, l_rm, l_rv, wbc, act, pre)
        ctx.meta = (n, t, c, eps1, eps2)
        return kern, act

    @staticmethod
    def backward(ctx, gkern, gact):
        pooled, w1, g_w, g_b, g_rm, g_rv, w2, wa, l_w, l_b, l_rm, l_rv, wb, act, pre = ctx.saved_tensors
        n, t, c, eps1, eps2 = ctx.meta
        dev = pooled.device
        z = lambda ref: torch.zeros_like(ref)
        gkern = z(act.new_empty(n, 3, c)) if gkern is None else gkern.contiguous()
        gact = z(act) if gact is None else gact.contiguous()
        key = (dev, n, t, c)
        ws = _gate_ws.get(key)
        if ws is None:
            ws = _gate_ws[key] = torch.zeros(_lib.load().vitta_tam_gate_bwd_ws_floats(n, t, c), dtype=torch.float32,
                                             device=dev)
        e = lambda *sh: torch.empty(*sh, dtype=torch.float32, device=dev)
        gp, gw1, gw2, gwa, gwb = e(n * t, c), e(*w1.shape), e(*w2.shape), e(*wa.shape), e(*wb.shape)
        gb1 = e(2, g_w.shape[0])
        gb2 = e(2, l_w.shape[0])
        gpre, ghm = e(n * t, c // 4), e(n * t, c // 4)
        call("vitta_tam_gate_bwd", ptr(pooled), ptr(w1), _lib.make_bn(g_w, g_b, g_rm, g_rv, eps1), ptr(w2), ptr(wa),
             _lib.make_bn(l_w, l_b, l_rm, l_rv, eps2), ptr(wb), ptr(act), ptr(pre), ptr(gkern), ptr(gact), ptr(gp),
             ptr(gw1), ptr(gb1[0]), ptr(gb1[1]), ptr(gw2), ptr(gwa), ptr(gb2[0]), ptr(gb2[1]), ptr(gwb), ptr(gpre),
             ptr(ghm), ptr(ws), n, t, c, stream_ptr())
        return (gp, gw1, gb1[0], gb1[1], None, None, gw2, gwa, gb2[0], gb2[1], None, None, gwb, None, None, None)


# ----------------------------------------------------------------------------------------------
# BN-folded inference forward (the per-step clean evaluation, reference corpus/basics.py:691-713)
# ----------------------------------------------------------------------------------------------
class FoldedConvs:
    """Operands of the inference convolutions of one model: for every (convolution, eval-mode BatchNorm) pair the fp16
    hi/lo pieces of W' = k * W (k = gamma / sqrt(running_var + eps), per output channel) and the bias
    b' = beta - running_mean * k, so that  BN(conv(x, W)) = conv(x, W') + b'  costs no pass of its own: the GEMM epilogue
    adds b' (and the shortcut) and applies ReLU.  All layers are refreshed by TWO launches (vitta_split_multi with the
    fold pointers + vitta_bn_fold_bias_multi) whenever a weight, a BatchNorm parameter or a running statistic changed."""

    def __init__(self, pairs):
        self.pairs = list(pairs)
        dev = self.pairs[0][0].weight.device
        self.slot = {id(conv): i for i, (conv, _) in enumerate(self.pairs)}
        self.bias = torch.empty(sum(bn.num_features for _, bn in self.pairs), dtype=torch.float32, device=dev)
        self.ops_, o = [], 0
        for conv, bn in self.pairs:
            n = conv.weight.numel()
            hi = torch.empty(n, dtype=torch.float16, device=dev)
            self.ops_.append((hi, torch.empty_like(hi), torch.zeros(1, dtype=torch.float32, device=dev),
                              self.bias[o:o + bn.num_features]))
            o += bn.num_features
        self._key = None
        self._stamp = None
        self._tables = None

    def _build(self, dev):
        blk = _lib.load().vitta_split_block_elems()
        arr = (_lib.VittaSplitTensor * len(self.pairs))()
        fb = (_lib.VittaFoldBias * len(self.pairs))()
        starts, b = [], 0
        for i, ((conv, bn), (hi, lo, am, bias)) in enumerate(zip(self.pairs, self.ops_)):
            r, t, c, tap_inner = _split_geom(conv.weight)
            e = arr[i]
            e.src, e.hi, e.lo, e.amax = conv.weight.data_ptr(), hi.data_ptr(), lo.data_ptr(), am.data_ptr()
            e.R, e.T, e.Cc, e.mode, e.src_tap_inner, e.compute_amax, e.n = r, t, c, 0, tap_inner, 1, r * t * c
            e.fold_w, e.fold_rv, e.fold_eps = bn.weight.data_ptr(), bn.running_var.data_ptr(), float(bn.eps)
            starts.append(b)
            b += (e.n + blk - 1) // blk
            f = fb[i]
            f.w, f.b, f.rm, f.rv = (bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                    bn.running_var.data_ptr())
            f.out, f.eps, f.C = bias.data_ptr(), float(bn.eps), bn.num_features
        host = (pinned_bytes(arr), torch.tensor(starts, dtype=torch.int32).pin_memory(), pinned_bytes(fb))
        self._tables = (host, tuple(h.to(dev, non_blocking=True) for h in host), len(self.pairs), b)

    def refresh(self):
        dev = self.bias.device
        key = tuple((conv.weight.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                     bn.running_var.data_ptr()) for conv, bn in self.pairs)
        stamp = (_weight_epoch,) + tuple((conv.weight._version, bn.weight._version, bn.bias._version,
                                          bn.running_mean._version, bn.running_var._version) for conv, bn in self.pairs)
        if key != self._key:
            if any(_split_geom(conv.weight) is None for conv, _ in self.pairs):
                raise _lib.VittaError("FoldedConvs: convolution weights must be contiguous")
            self._build(dev)
            self._key, self._stamp = key, None
        if stamp == self._stamp:
            return
        _, (t, starts, fb), n, blocks = self._tables
        call("vitta_split_multi", ptr(t), ptr(starts), n, blocks, 1, stream_ptr())
        call("vitta_bn_fold_bias_multi", ptr(fb), n, stream_ptr())
        self._stamp = stamp

    def get(self, conv):
        return self.ops_[self.slot[id(conv)]]


def conv2d_folded(x, conv, folds, relu, residual=None):
    """[relu](BN(conv(x)) [+ residual]) as ONE launch; x channels_last, inference only (no autograd)."""
    _require_cuda(x, "conv2d_folded")
    if not x.is_contiguous(memory_format=CL):
        x = x.contiguous(memory_format=CL)
    hi, lo, am, bias = folds.get(conv)
    cout, cin, kh, kw = conv.weight.shape
    stride, pad = conv.stride[0], conv.padding[0]
    f, _, h, w = x.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    y = torch.empty((f, cout, ho, wo), dtype=torch.float32, device=x.device, memory_format=CL)
    if residual is not None and not (residual.shape == y.shape and residual.is_contiguous(memory_format=CL)):
        raise _lib.VittaError("conv2d_folded: residual must be channels_last with the shape of the result")
    y_am = new_amax(x.device)
    call("vitta_conv2d_f16x3_infer", ptr(x), ptr(operand_amax(x)), f, h, w, cin, ptr(hi), ptr(lo), ptr(am), cout, kh, kw,
         stride, pad, ptr(y), ptr(bias), ptr(residual), 1 if relu else 0, ptr(y_am), stream_ptr())
    _attach_amax(y, y_am)
    return y


def frame_mean_cl(x):
    """(F, C, H, W) channels_last -> (F, C) spatial means (vitta_frame_mean), inference helper."""
    f, c, h, w = x.shape
    out = torch.empty(f, c, dtype=torch.float32, device=x.device)
    call("vitta_frame_mean", ptr(x), f, h * w, c, ptr(out), stream_ptr())
    return out
