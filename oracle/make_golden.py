"""TEST INFRASTRUCTURE -- generate ``tests/golden/*.npz`` by running the UNMODIFIED reference on CPU.

Run in the build container (needs /root/reference):   python -m oracle.make_golden [case ...]

The reference has no tests or golden vectors, so the parity pin is built here:
the reference's own ``compute_statistics`` and ``tta_standard`` (corpus/basics.py:220-307, 403-747) are
executed end to end on seeded synthetic data (``vitta_b200.synth``), with recording wrappers around
its hook class, its model and ``compute_pred_consis``.  Only *outputs* are stored; inputs and weights
are regenerated from the same seeds by the tests.

Deviations from the shipped scripts, all via ``args`` (no reference source is modified):
  * Dropout(0.8)/Dropout(0.5)/DropPath are set to p=0 so the train-mode forward is deterministic;
  * ``lr`` is raised (1e-3) so three SGD steps move the weights by a testable amount;
  * dataset factories are replaced by synthetic tensor datasets; ``np.save`` of the ragged stat list
    (basics.py:306-307, broken on numpy>=1.24) is intercepted to capture the lists.
"""
import copy
import logging
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

from oracle import ref_harness
from vitta_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TANET_CASES = {
    # name: dict(config)
    "tanet_t8_r64_consis_l1": dict(K=101, T=8, N=2, M=2, res=64, reg_type="l1_loss", consis=True, steps=3,
                                   lr=1e-3, moving_avg=True),
    "tanet_t8_r64_stats_mse": dict(K=101, T=8, N=2, M=1, res=64, reg_type="mse_loss", consis=False, steps=2,
                                   lr=1e-3, moving_avg=True),
    "tanet_t16_r224_stats_l1": dict(K=101, T=16, N=1, M=1, res=224, reg_type="l1_loss", consis=False, steps=1,
                                    lr=1e-3, moving_avg=True),
    # option rows of SURVEY 8(f) rank 4 at model level: KLD against running MEANS of the statistics (moving_avg=False,
    # AverageMeterTensor), and --update_only_bn_affine (everything frozen except norm affine parameters, Adam)
    "tanet_t8_r64_stats_kld_avg": dict(K=101, T=8, N=2, M=1, res=64, reg_type="kld", consis=False, steps=3,
                                       lr=1e-6, moving_avg=False),   # KLD sums over channels: large gradients
    "tanet_t8_r64_consis_l1_bnaffine": dict(K=101, T=8, N=2, M=2, res=64, reg_type="l1_loss", consis=True, steps=2,
                                            lr=1e-3, moving_avg=True, bn_affine=True),
    # tta_standard mode (corpus/basics.py:414-419,519-530): a fresh model copy, optimiser and hooks for every batch,
    # momentum_mvg = 1 (no accumulation of target statistics), several gradient steps on the same batch
    # --stat_reg BNS (utils/BNS_utils.py:19-77): statistics of every BN *input* (BatchNorm1d of the TAM branches
    # included) against that layer's running statistics, EMA from zeros (running_manner)
    "tanet_t8_r64_bns_l1": dict(K=101, T=8, N=2, M=2, res=64, reg_type="l1_loss", consis=True, steps=2,
                                lr=1e-3, moving_avg=True, stat_reg="BNS"),
    # --before_norm: statistics (source and target) of the norm layers' INPUT instead of their output
    "tanet_t8_r64_stats_l1_before_norm": dict(K=101, T=8, N=2, M=1, res=64, reg_type="l1_loss", consis=False, steps=2,
                                              lr=1e-3, moving_avg=True, before_norm=True),
    "tanet_t8_r64_standard_l1": dict(K=101, T=8, N=2, M=2, res=64, reg_type="l1_loss", consis=True, steps=2,
                                     lr=1e-3, moving_avg=True, mode="tta_standard", momentum_mvg=1.0, gsteps=2),
}

SWIN_CASES = {
    "swin_tiny_t16_r112_consis_l1": dict(K=101, T=16, N=1, M=2, res=112, embed_dim=64, depths=[2, 2], heads=[2, 4],
                                         window=(8, 7, 7), reg_type="l1_loss", consis=True, steps=2, lr=1e-3,
                                         chosen=["module.backbone.layers.1", "module.backbone.norm"],
                                         momentum_mvg=0.05, lambda_consis=0.05),
    "swin_tiny_t32_r56_stats_l1": dict(K=101, T=32, N=2, M=1, res=56, embed_dim=32, depths=[2, 2, 2], heads=[1, 2, 4],
                                       window=(8, 7, 7), reg_type="l1_loss", consis=False, steps=2, lr=1e-3,
                                       chosen=["module.backbone.layers.1", "module.backbone.layers.2",
                                               "module.backbone.norm"],
                                       momentum_mvg=0.05, lambda_consis=0.05, sample_views=False),
}

# option rows at model level for Video-Swin; kept apart from SWIN_CASES so that the default GPU suite
# (which iterates SWIN_CASES) only contains cases that have run on hardware
SWIN_OPTION_CASES = {
    # --update_only_bn_affine on Swin: everything frozen except the LayerNorm affine parameters, Adam (basics.py:552-557)
    "swin_tiny_t16_r112_consis_l1_lnaffine": dict(K=101, T=16, N=1, M=2, res=112, embed_dim=64, depths=[2, 2], heads=[2, 4],
                                                  window=(8, 7, 7), reg_type="l1_loss", consis=True, steps=2, lr=1e-3,
                                                  chosen=["module.backbone.layers.1", "module.backbone.norm"],
                                                  momentum_mvg=0.05, lambda_consis=0.05, bn_affine=True),
    # MSE alignment against running MEANS of the statistics (moving_avg=False, AverageMeterTensor), one view, no consistency
    "swin_tiny_t32_r56_stats_mse_avg": dict(K=101, T=32, N=2, M=1, res=56, embed_dim=32, depths=[2, 2, 2], heads=[1, 2, 4],
                                            window=(8, 7, 7), reg_type="mse_loss", consis=False, steps=3, lr=1e-3,
                                            chosen=["module.backbone.layers.1", "module.backbone.layers.2",
                                                    "module.backbone.norm"],
                                            momentum_mvg=0.05, lambda_consis=0.05, sample_views=False, moving_avg=False),
}


class _ListDataset(torch.utils.data.Dataset):
    def __init__(self, x, y):
        self.x, self.y = x, y

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i]


class _NpProxy:
    """numpy stand-in inside corpus.basics: intercept np.save of the ragged stat lists."""

    def __init__(self, sink):
        self._sink = sink

    def __getattr__(self, name):
        return getattr(np, name)

    def save(self, path, obj, allow_pickle=True):
        self._sink[os.path.basename(path)] = [np.asarray(o) for o in obj]


class _CopyProxy:
    def __init__(self, sink):
        self._sink = sink

    def __getattr__(self, name):
        return getattr(copy, name)

    on_model_copy = None   # callback: tta_standard mode re-creates model + hooks per batch

    def deepcopy(self, obj, *a):
        out = copy.deepcopy(obj, *a)
        if isinstance(obj, nn.Module):
            self._sink.append(out)
            if self.on_model_copy is not None:
                self.on_model_copy()
        return out


def _base_args(ref, cfg, arch):
    args = ref["utils.opts"].parser.parse_args([])
    args.evaluate_baselines, args.baseline = False, "source"
    args.arch, args.dataset, args.gpus = arch, "ucf101", [0]
    args.num_classes = cfg["K"]
    args.batch_size, args.clip_length, args.workers = cfg["N"], cfg["T"], 0
    args.sample_style, args.test_crops, args.num_clips = "uniform-1", 1, 1
    args.verbose = False
    args.reg_type, args.moving_avg = cfg["reg_type"], cfg.get("moving_avg", True)
    args.if_pred_consistency = cfg["consis"]
    args.if_sample_tta_aug_views = cfg.get("sample_views", True)
    args.n_augmented_views = cfg["M"]
    args.lr = cfg["lr"]
    args.update_only_bn_affine = cfg.get("bn_affine", False)
    args.stat_reg = cfg.get("stat_reg", "mean_var")
    args.before_norm = cfg.get("before_norm", False)
    args.if_tta_standard = cfg.get("mode", "tta_online")
    args.n_gradient_steps = cfg.get("gsteps", 1)
    args.momentum_mvg = cfg.get("momentum_mvg", 0.1)
    args.lambda_pred_consis = cfg.get("lambda_consis", 0.1)
    args.stat_type = ["spatiotemp"]
    if "chosen" in cfg:
        args.chosen_blocks = cfg["chosen"]
    args.result_dir = tempfile.mkdtemp(prefix="vitta_golden_")
    return args


def _tanet_inputs(cfg, n_batches, tag, seed):
    """n_batches loader batches in TANet layout (N, M*T*3, H, W)."""
    out = []
    for b in range(n_batches):
        v = synth.synth_video(cfg["N"], cfg["M"] if tag == "tta" else 1, cfg["T"], cfg["res"], seed=seed + b,
                              gauss_sigma=0.38 if tag != "clean" else 0.0, tag=tag)
        out.append(synth.tanet_loader_tensor(v))
    return out


def _swin_inputs(cfg, n_batches, tag, seed):
    out = []
    for b in range(n_batches):
        v = synth.synth_video(cfg["N"], cfg["M"] if tag == "tta" else 1, cfg["T"], cfg["res"], seed=seed + b,
                              gauss_sigma=0.38 if tag != "clean" else 0.0, tag=tag)
        out.append(synth.swin_loader_tensor(v))
    return out


def run_model_case(name, cfg, arch):
    ref = ref_harness.load_reference()
    basics = ref["corpus.basics"]
    nsu = ref["utils.norm_stats_utils"]
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    if arch == "tanet":
        model = ref_harness.build_reference_tsn(cfg["K"], cfg["T"])
        model.base_model.fc.p = 0.0  # Dropout(0.8) -> deterministic
        make_inputs = _tanet_inputs
    else:
        model = ref_harness.build_reference_swin(cfg["K"], (2, 4, 4), cfg["window"], 0.0, cfg["embed_dim"],
                                                 cfg["depths"], cfg["heads"])
        model.cls_head.dropout.p = 0.0
        make_inputs = _swin_inputs
    sd = synth.synth_state_dict(model.state_dict(), seed=1)
    model.load_state_dict(sd, strict=True)
    if arch != "tanet":
        model = torch.nn.DataParallel(model, device_ids=[0])  # keeps the 'module.' prefix chosen_blocks need
    args = _base_args(ref, cfg, arch)

    # ---- source statistics through the reference's own compute_statistics ------------------
    saved = {}
    basics.np = _NpProxy(saved)
    clean = make_inputs(cfg, 2, "clean", seed=100)
    labels = synth.synth_labels(cfg["N"], cfg["K"], seed=3)
    clean_ds = _ListDataset(torch.cat(clean, 0), labels.repeat(2))
    basics.get_dataset_tanet = lambda a, split=None, dataset_type=None: clean_ds
    basics.get_dataset_videoswin = lambda a, split=None, dataset_type=None: clean_ds
    a2 = copy.copy(args)
    a2.stat_type = "spatiotemp"
    a2.before_norm = cfg.get("before_norm", False)
    basics.compute_statistics(model, args=a2, log_time="x")
    src_mean = saved["list_spatiotemp_mean_x.npy"]
    src_var = saved["list_spatiotemp_var_x.npy"]
    basics.np = np

    def _obj(lst):
        arr = np.empty(len(lst), dtype=object)
        for i, v in enumerate(lst):
            arr[i] = v
        return arr
    args.spatiotemp_mean_clean_file = os.path.join(args.result_dir, "m.npy")
    args.spatiotemp_var_clean_file = os.path.join(args.result_dir, "v.npy")
    np.save(args.spatiotemp_mean_clean_file, _obj(src_mean), allow_pickle=True)
    np.save(args.spatiotemp_var_clean_file, _obj(src_var), allow_pickle=True)

    # ---- the adaptation loop through the reference's own tta_standard ----------------------
    steps = cfg["steps"]
    tta_in = make_inputs(cfg, steps, "tta", seed=200)
    eval_in = make_inputs(cfg, steps, "eval", seed=300)
    tta_ds = _ListDataset(torch.cat(tta_in, 0), labels.repeat(steps))
    eval_ds = _ListDataset(torch.cat(eval_in, 0), labels.repeat(steps))

    def fake_ds(a, split=None, dataset_type=None):
        return tta_ds if dataset_type == "tta" else eval_ds
    basics.get_dataset_tanet = fake_ds
    basics.get_dataset_videoswin = fake_ds

    rec = {"hooks": [], "outputs": [], "consis": [], "models": []}
    base_cls = nsu.CombineNormStatsRegHook_onereg

    class Recording(base_cls):
        _count = 0

        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.idx = Recording._count
            Recording._count += 1

        def hook_fn(self, module, inp, out):
            super().hook_fn(module, inp, out)
            if isinstance(module, nn.BatchNorm1d):
                rec["hooks"].append((self.idx, None, None, 0.0))
            else:
                rec["hooks"].append((self.idx, self.mean_avgmeter_spatiotemp.avg.detach().numpy().copy(),
                                     self.var_avgmeter_spatiotemp.avg.detach().numpy().copy(),
                                     float(self.r_feature.detach())))
    nsu.CombineNormStatsRegHook_onereg = Recording

    base_bns = basics.BNFeatureHook

    class RecordingBNS(base_bns):
        """--stat_reg BNS: the hooks are BNFeatureHook (utils/BNS_utils.py:19-77), also on the TAM's BatchNorm1d."""

        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.idx = Recording._count
            Recording._count += 1

        def hook_fn(self, module, inp, out):
            super().hook_fn(module, inp, out)
            rec["hooks"].append((self.idx, self.mean.detach().numpy().copy(), self.var.detach().numpy().copy(),
                                 float(self.r_feature.detach())))
    basics.BNFeatureHook = RecordingBNS

    def out_hook(m, i, o):
        rec["outputs"].append(o)
    top = model
    top.register_forward_hook(out_hook)
    orig_consis = basics.compute_pred_consis

    def consis_rec(p):
        v = orig_consis(p)
        rec["consis"].append(float(v.detach()))
        return v
    basics.compute_pred_consis = consis_rec
    basics.cp = _CopyProxy(rec["models"])

    def _new_model():
        Recording._count = 0      # hooks are re-created for every model copy: index them per copy
    basics.cp.on_model_copy = _new_model
    try:
        top1 = basics.tta_standard(model, nn.CrossEntropyLoss(), args=args, logger=logging.getLogger("golden"),
                                   writer=None)
    finally:
        basics.BNFeatureHook = base_bns
        nsu.CombineNormStatsRegHook_onereg = base_cls
        basics.compute_pred_consis = orig_consis
        basics.cp = copy

    adapted = rec["models"][-1]   # tta_online: the one copy; tta_standard: the copy adapted on the last batch
    n_hooks = Recording._count
    gsteps = cfg.get("gsteps", 1)
    out = {"top1": np.float32(top1[0]), "n_hooks": np.int64(n_hooks), "steps": np.int64(steps),
           "gsteps": np.int64(gsteps)}
    for i, (m, v) in enumerate(zip(src_mean, src_var)):
        out["src_mean/%d" % i] = m.astype(np.float32)
        out["src_var/%d" % i] = v.astype(np.float32)
    assert len(rec["hooks"]) == n_hooks * steps * gsteps, (len(rec["hooks"]), n_hooks, steps, gsteps)
    assert len(rec["outputs"]) == steps * (gsteps + 1)
    for s in range(steps):
        for j in range(gsteps):
            # key prefix: "step<s>" (one gradient step per batch) or "step<s>.<j>" (n_gradient_steps > 1)
            pre = "step%d" % s if gsteps == 1 else "step%d.%d" % (s, j)
            fw = s * gsteps + j          # index of this training forward
            rsum = 0.0
            # hooks fire in execution order (bn1, tam.G.1, tam.L.1, bn2, ...); index them by creation order
            step_recs = sorted(rec["hooks"][fw * n_hooks:(fw + 1) * n_hooks], key=lambda t: t[0])
            for h in range(n_hooks):
                idx, em, ev, r = step_recs[h]
                assert idx == h
                if em is not None:
                    out["%s/ema_mean/%d" % (pre, h)] = em
                    out["%s/ema_var/%d" % (pre, h)] = ev
                out["%s/r_feature/%d" % (pre, h)] = np.float32(r)
                rsum += r
            out["%s/loss_reg" % pre] = np.float32(rsum)
            if cfg["consis"]:
                out["%s/loss_consis" % pre] = np.float32(rec["consis"][fw])
            tr = rec["outputs"][s * (gsteps + 1) + j]
            out["%s/train_logits" % pre] = (tr if arch == "tanet" else tr[1]).detach().numpy()   # Swin: (N, V, K)
        ev_ = rec["outputs"][s * (gsteps + 1) + gsteps]
        out["step%d/eval_logits" % s] = (ev_ if arch == "tanet" else ev_[0]).detach().numpy()
    # weight movement after all steps: per-tensor delta norms + a few full deltas
    new_sd = adapted.state_dict()
    names, dn = [], []
    for k, v in new_sd.items():
        k0 = k[len("module."):] if k.startswith("module.") else k
        if not v.is_floating_point():
            continue
        d = (v.detach() - sd[k0]).double()
        names.append(k0)
        dn.append([float(d.norm()), float(d.sum())])
    out["delta_names"] = np.array(names)
    out["delta_norm_sum"] = np.asarray(dn, np.float64)
    keep = ([n for n in names if n.endswith("conv1.weight") and "layer" not in n] +
            [n for n in names if "layer3.1.net.bn2" in n and (n.endswith("weight") or n.endswith("bias"))] +
            [n for n in names if "layer4.2.net.conv3.weight" in n] +
            [n for n in names if n.startswith("new_fc") or n.startswith("cls_head")] +
            [n for n in names if "layers.1.blocks.1.attn.relative_position_bias_table" in n] +
            [n for n in names if "layers.1.blocks.1.attn.qkv.bias" in n] +
            [n for n in names if "layers.1.blocks.0.norm2" in n])
    for n in keep:
        k = n if n in new_sd else "module." + n
        d = (new_sd[k].detach() - sd[n]).reshape(-1)
        out["delta/" + n] = d[:4096].numpy().copy()   # leading slice only: keeps the fixtures small
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print("wrote", name, "hooks", n_hooks, "loss_reg", [float(v) for k, v in sorted(out.items()) if k.endswith("/loss_reg")],
          "consis", rec["consis"], "top1", float(top1[0]))


def run_unit_case():
    """Hook / loss / TAM operators of the reference on small random tensors, incl. autograd grads."""
    ref = ref_harness.load_reference()
    nsu = ref["utils.norm_stats_utils"]
    bns = ref["utils.BNS_utils"]
    pcu = ref["utils.pred_consistency_utils"]
    tm = ref["models.tanet_models.temporal_module"]
    g = np.random.Generator(np.random.PCG64(7))
    out = {}

    def rnd(*shape, scale=1.0, shift=0.0):
        return torch.from_numpy((g.normal(size=shape) * scale + shift).astype(np.float32))

    # --- CombineNormStatsRegHook_onereg on BN2d / BN3d / LN outputs, 3 steps, all reg types, both meters
    shapes = {"bn2d": (2 * 4, 6, 5, 7), "bn3d": (2, 6, 4, 5, 7), "ln": (2, 4, 5, 7, 6)}
    for kind, shp in shapes.items():
        c = 6
        for reg in ("l1_loss", "mse_loss", "kld"):
            for mavg in (True, False):
                mod = {"bn2d": nn.BatchNorm2d(c), "bn3d": nn.BatchNorm3d(c), "ln": nn.LayerNorm(c)}[kind]
                mod.eval()
                src_m = (g.normal(size=c) * 0.3).astype(np.float32)
                src_v = g.uniform(0.5, 1.5, size=c).astype(np.float32)
                hook = nsu.CombineNormStatsRegHook_onereg(
                    mod, clip_len=4, spatiotemp_stats_clean_tuple=(src_m, src_v), reg_type=reg, moving_avg=mavg,
                    momentum=0.1 if reg != "kld" else 0.9, stat_type_list=["spatiotemp"], reduce_dim=True,
                    before_norm=False, if_sample_tta_aug_views=True, n_augmented_views=2)
                key = "hook/%s/%s/%d" % (kind, reg, int(mavg))
                out[key + "/src_mean"], out[key + "/src_var"] = src_m, src_v
                for s in range(3):
                    feat = rnd(*shp, scale=1.3, shift=0.2).requires_grad_(True)
                    hook.hook_fn(mod, (feat,), feat * 1.0)  # feed `feat*1` as the layer output
                    # the graph is output->feat, so feat.grad == dL/d(output)
                    hook.r_feature.backward()
                    out["%s/s%d/feat" % (key, s)] = feat.detach().numpy()
                    out["%s/s%d/r" % (key, s)] = np.float32(hook.r_feature.detach())
                    out["%s/s%d/grad" % (key, s)] = feat.grad.numpy().copy()
                    out["%s/s%d/ema_mean" % (key, s)] = hook.mean_avgmeter_spatiotemp.avg.detach().numpy().copy()
                    out["%s/s%d/ema_var" % (key, s)] = hook.var_avgmeter_spatiotemp.avg.detach().numpy().copy()
    # --- ComputeNormStatsHook, all stat types
    for kind, shp in shapes.items():
        for st in ("spatiotemp", "temp", "temp_v2", "spatial"):
            mod = {"bn2d": nn.BatchNorm2d(6), "bn3d": nn.BatchNorm3d(6), "ln": nn.LayerNorm(6)}[kind]
            hook = nsu.ComputeNormStatsHook(mod, clip_len=4, stat_type=st, before_norm=False, batch_size=2)
            feat = rnd(*shp, scale=0.7, shift=-0.4)
            hook.hook_fn(mod, (feat,), feat)
            key = "stat/%s/%s" % (kind, st)
            out[key + "/feat"] = feat.numpy()
            out[key + "/mean"] = hook.batch_mean.numpy()
            out[key + "/var"] = hook.batch_var.numpy()
    # --- BNFeatureHook (stat_reg BNS)
    for running in (True, False):
        mod = nn.BatchNorm2d(6).eval()
        mod.running_mean.copy_(rnd(6, scale=0.2))
        mod.running_var.copy_(torch.from_numpy(g.uniform(0.5, 1.5, 6).astype(np.float32)))
        hook = bns.BNFeatureHook(mod, reg_type="l1_loss", running_manner=running, use_src_stat_in_reg=True,
                                 momentum=0.1)
        key = "bns/%d" % int(running)
        out[key + "/running_mean"] = mod.running_mean.numpy().copy()
        out[key + "/running_var"] = mod.running_var.numpy().copy()
        for s in range(2):
            x = rnd(8, 6, 5, 7, scale=1.1).requires_grad_(True)
            hook.hook_fn(mod, (x,), None)
            hook.r_feature.backward()
            out["%s/s%d/x" % (key, s)] = x.detach().numpy()
            out["%s/s%d/r" % (key, s)] = np.float32(hook.r_feature.detach())
            out["%s/s%d/grad" % (key, s)] = x.grad.numpy().copy()
    # --- compute_pred_consis with grads
    for (b, v, k) in ((2, 2, 101), (3, 4, 17), (1, 2, 400)):
        p = rnd(b, v, k, scale=2.0).requires_grad_(True)
        loss = pcu.compute_pred_consis(p)
        loss.backward()
        key = "consis/%d_%d_%d" % (b, v, k)
        out[key + "/preds"] = p.detach().numpy()
        out[key + "/loss"] = np.float32(loss.detach())
        out[key + "/grad"] = p.grad.numpy().copy()
    # --- TAM forward/backward (eval-mode BN1d, as in the adaptation forward)
    import contextlib
    import io
    for (n, t, c, h, w) in ((2, 8, 16, 5, 7), (1, 16, 32, 7, 7)):
        with contextlib.redirect_stdout(io.StringIO()):
            tam = tm.TAM(c, t)
        tsd = synth.synth_state_dict(tam.state_dict(), seed=5)
        tam.load_state_dict(tsd)
        tam.train()
        for m in tam.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.eval()
        x = rnd(n * t, c, h, w).abs().requires_grad_(True)
        y = tam(x)
        go = rnd(*y.shape)
        y.backward(go)
        key = "tam/%d_%d_%d_%d_%d" % (n, t, c, h, w)
        out[key + "/x"], out[key + "/y"], out[key + "/go"] = x.detach().numpy(), y.detach().numpy(), go.numpy()
        out[key + "/gx"] = x.grad.numpy().copy()
        for pn, p in tam.named_parameters():
            out["%s/gp/%s" % (key, pn)] = p.grad.numpy().copy()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "units.npz"), **out)
    print("wrote units", len(out), "arrays")


def run_views_case():
    """Frame indices of the reference's own view sampler (models/tanet_models/video_dataset.py:159-196) for the
    deterministic styles, over a grid of video lengths / clip lengths / view counts."""
    import importlib
    ref_harness.load_reference()
    vd = importlib.import_module("models.tanet_models.video_dataset")

    class Rec:
        def __init__(self, n):
            self.num_frames = n
    out = {}
    for style in ("uniform", "dense", "uniform_equidist", "dense_equidist"):
        for nf in (9, 16, 31, 64, 100, 177, 300):
            for t in (8, 16, 32):
                for views in (1, 2, 3):
                    ds = object.__new__(vd.Video_TANetDataSet)
                    ds.num_segments, ds.new_length, ds.n_tta_aug_views = t, 1, views
                    idx = ds._sample_tta_augmented_views(Rec(nf), style)
                    idx = np.minimum(np.asarray(idx), nf - 1)          # the clamp of get() (:328)
                    out["%s/%d/%d/%d" % (style, nf, t, views)] = idx.astype(np.int64)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "views.npz"), **out)
    print("wrote views", len(out), "index vectors")
    # evaluation clips: _get_test_indices for --sample_style uniform-N / dense-N
    tst = {}
    for style in ("uniform-1", "uniform-3", "dense-1", "dense-2", "dense-4"):
        for nf in (5, 9, 16, 31, 64, 100, 177, 300):
            for t in (8, 16, 32):
                ds = object.__new__(vd.Video_TANetDataSet)
                ds.num_segments, ds.new_length, ds.test_sample = t, 1, style
                tst["%s/%d/%d" % (style, nf, t)] = np.minimum(np.asarray(ds._get_test_indices(Rec(nf))), nf - 1).astype(np.int64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "test_indices.npz"), **tst)
    print("wrote test_indices", len(tst), "index vectors")
    # the random styles (one view per call) under a seeded numpy legacy generator
    rnd = {}
    for style in ("uniform_rand", "dense_rand", "random"):
        for nf in (5, 9, 16, 31, 64, 100, 177, 300):
            for t in (8, 16, 32):
                ds = object.__new__(vd.Video_TANetDataSet)
                ds.num_segments, ds.new_length, ds.n_tta_aug_views = t, 1, 1
                np.random.seed(nf * 100 + t)
                a = np.asarray(ds._sample_tta_augmented_views(Rec(nf), style))
                b = np.asarray(ds._sample_tta_augmented_views(Rec(nf), style))       # second draw from the same stream
                rnd["%s/%d/%d" % (style, nf, t)] = np.minimum(np.stack([a, b]), nf - 1).astype(np.int64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "views_rand.npz"), **rnd)
    print("wrote views_rand", len(rnd), "index arrays")
    # the Video-Swin loader's clean evaluation clip (SampleFrames.get_seq_frames, test mode)
    tb = importlib.import_module("models.videoswintransformer_models.transforms_backup")
    seq = {}
    for nf in (1, 2, 9, 16, 17, 31, 33, 64, 100, 177, 300):
        for t in (8, 16, 32):
            sf = object.__new__(tb.SampleFrames)
            sf.clip_len, sf.test_mode = t, True
            seq["%d/%d" % (nf, t)] = np.minimum(np.asarray(sf.get_seq_frames(nf)), nf - 1).astype(np.int64)
    # RandomResizedCrop.get_crop_bbox under seeded numpy + random generators (8 consecutive boxes per frame size)
    import random
    for ih, iw in ((256, 341), (256, 455), (256, 256), (341, 256), (64, 85), (120, 30)):
        np.random.seed(ih * 7 + iw)
        random.seed(ih * 7 + iw)
        seq["bbox/%d/%d" % (ih, iw)] = np.asarray(
            [tb.RandomResizedCrop.get_crop_bbox((ih, iw), (0.08, 1.0), (3 / 4, 4 / 3)) for _ in range(8)], np.int64)
    # draw ORDER of the TTA pipeline tail: RandomResizedCrop.__call__ then Flip.__call__ (flip_ratio 0) of the reference on
    # two consecutive items; the generators' next values afterwards fix the stream positions.  (__init__ bypassed: it only
    # validates with mmcv.is_tuple_of, and mmcv is absent.)
    rrc = object.__new__(tb.RandomResizedCrop)
    rrc.area_range, rrc.aspect_ratio_range, rrc.lazy = (0.08, 1.0), (3 / 4, 4 / 3), False
    fl = object.__new__(tb.Flip)
    fl.flip_ratio, fl.direction, fl.flip_label_map, fl.left_kp, fl.right_kp, fl.lazy = 0, 'horizontal', None, None, None, False
    np.random.seed(3)
    random.seed(3)
    boxes = []
    for h, w in ((40, 53), (55, 40)):                     # = swin_rescale_size of 64x48 and 44x60 frames to short edge 40
        res = {"imgs": [np.zeros((h, w, 3), np.uint8)] * 2, "img_shape": (h, w), "modality": "RGB"}
        res = fl(rrc(res))
        assert not res["flip"]
        boxes.append(res["crop_bbox"])
    seq["pipeline/boxes"] = np.asarray(boxes, np.int64)
    seq["pipeline/next"] = np.asarray([np.random.rand(), random.random()], np.float64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "swin_seq.npz"), **seq)
    print("wrote swin_seq", len(seq), "index vectors")


def run_crops_case():
    """Per-view random multi-scale crops of the reference's own transform (models/tanet_models/transforms.py:277-384):
    (a) crop boxes of ``_sample_crop_size`` for seeded ``random`` streams over a grid of frame / input sizes,
    (b) the whole transform (``crop`` + ``resize(BILINEAR)`` through the installed Pillow) on seeded uint8 frames."""
    import importlib
    import random

    from PIL import Image
    ref_harness.load_reference()
    tr = importlib.import_module("models.tanet_models.transforms")
    out = {}
    for iw, ih, inp in ((320, 240, 224), (340, 256, 224), (455, 256, 224), (171, 128, 112), (224, 224, 224),
                        (256, 256, 224), (64, 48, 32), (398, 224, 224)):
        op = tr.SubgroupWise_MultiScaleCrop_TANet(input_size=inp, n_temp_clips=2, clip_len=4)
        random.seed(1000 + iw)
        out["boxes/%d/%d/%d/%d" % (iw, ih, inp, 1000 + iw)] = np.asarray(
            [op._sample_crop_size((iw, ih)) for _ in range(64)], np.int64)
    rng = np.random.Generator(np.random.PCG64(11))
    for name, (ih, iw, inp, views, t, seed) in {"a": (48, 64, 32, 2, 3, 5), "b": (60, 44, 32, 3, 2, 6),
                                                "c": (33, 80, 24, 2, 2, 7)}.items():
        frames = rng.integers(0, 256, (views * t, ih, iw, 3), dtype=np.uint8)
        op = tr.SubgroupWise_MultiScaleCrop_TANet(input_size=inp, n_temp_clips=views, clip_len=t)
        random.seed(seed)
        imgs, _ = op(([Image.fromarray(f) for f in frames], 0))
        random.seed(seed)
        boxes = [op._sample_crop_size((iw, ih)) for _ in range(views)]      # the same draws, recorded
        out["xf/%s/frames" % name] = frames
        out["xf/%s/meta" % name] = np.asarray([inp, views, t, seed], np.int64)
        out["xf/%s/boxes" % name] = np.asarray(boxes, np.int64)
        out["xf/%s/out" % name] = np.stack([np.asarray(im) for im in imgs])
    # (c) the other spatial path: GroupScale_TANet + GroupCenterCrop_TANet (corpus/basics.py:1259-1263)
    for name, (ih, iw, z, inp) in {"a": (48, 64, 40, 32), "b": (64, 48, 40, 32), "c": (40, 40, 40, 32),
                                   "d": (51, 77, 36, 33), "e": (30, 100, 64, 48)}.items():
        frames = rng.integers(0, 256, (2, ih, iw, 3), dtype=np.uint8)
        imgs, _ = tr.GroupCenterCrop_TANet(inp)(tr.GroupScale_TANet(z)(([Image.fromarray(f) for f in frames], 0)))
        out["sc/%s/frames" % name] = frames
        out["sc/%s/meta" % name] = np.asarray([z, inp], np.int64)
        out["sc/%s/out" % name] = np.stack([np.asarray(im) for im in imgs])
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "crops.npz"), **out)
    print("wrote crops", len(out), "arrays")


def run_loader_case():
    """Whole items of the reference's own TANet loader: ``corpus.basics.get_dataset_tanet`` -> ``Video_TANetDataSet.
    __getitem__`` (index sampling, PIL conversion, per-view random multi-scale crop or scale + centre crop, Stack,
    ToTorchFormatTensor, GroupNormalize), unmodified.  Only the decoder is replaced: ``decord.VideoReader`` is an
    in-memory reader over seeded uint8 frames (decoding is out of scope)."""
    import importlib
    import random
    import tempfile
    ref = ref_harness.load_reference()
    vd = importlib.import_module("models.tanet_models.video_dataset")
    basics = importlib.import_module("corpus.basics")
    rng = np.random.Generator(np.random.PCG64(23))
    videos = {"v0": rng.integers(0, 256, (11, 48, 64, 3), dtype=np.uint8),
              "v1": rng.integers(0, 256, (20, 60, 44, 3), dtype=np.uint8)}
    labels = {"v0": 3, "v1": 7}

    class Batch:
        def __init__(self, a):
            self.a = a

        def asnumpy(self):
            return self.a

    class Reader:
        def __init__(self, path):
            self.frames = videos[os.path.basename(path)[:-4]]
            self._num_frame = len(self.frames)

        def get_batch(self, idx):
            return Batch(self.frames[np.asarray(idx)])
    vd.decord.VideoReader = Reader
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        lst = os.path.join(tmp, "list.txt")
        with open(lst, "w") as f:
            for k, v in videos.items():
                f.write("%s %d %d\n" % (k, len(v), labels[k]))
        for case, (kind, views, rand_crop, seed) in {"tta_randcrop": ("tta", 2, True, 31), "tta_center": ("tta", 2, False, 32),
                                                    "eval": ("eval", 2, True, 33), "tta_3views": ("tta", 3, True, 34),
                                                    "tta_3crops": ("tta", 2, True, 35)}.items():
            args = ref["utils.opts"].parser.parse_args([])
            args.arch, args.modality, args.vid_format = "tanet", "RGB", ".mp4"
            args.val_vid_list, args.video_data_dir = lst, tmp
            args.clip_length, args.input_size, args.scale_size, args.full_res = 4, 32, 40, False
            args.test_crops, args.sample_style, args.debug = (3 if case == "tta_3crops" else 1), "uniform-1", False
            args.if_sample_tta_aug_views, args.n_augmented_views = True, views
            args.if_spatial_rand_cropping = rand_crop
            args.tta_view_sample_style_list = ["uniform_equidist"]
            ds = basics.get_dataset_tanet(args, split="val", dataset_type=kind)
            random.seed(seed)
            for i, name in enumerate(videos):
                x, y = ds[i]
                out["%s/%s/x" % (case, name)] = x.numpy().astype(np.float32)
                out["%s/%s/y" % (case, name)] = np.asarray(y, np.int64)
            out["%s/meta" % case] = np.asarray([1 if kind == "tta" else 0, views, int(rand_crop), seed, args.test_crops],
                                               np.int64)
    for k, v in videos.items():
        out["video/%s" % k] = v
    np.savez_compressed(os.path.join(GOLDEN_DIR, "loader.npz"), **out)
    print("wrote loader", len(out), "arrays")


def main(argv):
    want = set(argv)
    if not want or "units" in want:
        run_unit_case()
    if not want or "views" in want:
        run_views_case()
    if not want or "crops" in want:
        run_crops_case()
    if not want or "loader" in want:
        run_loader_case()
    for name, cfg in TANET_CASES.items():
        if not want or name in want:
            run_model_case(name, cfg, "tanet")
    for name, cfg in list(SWIN_CASES.items()) + list(SWIN_OPTION_CASES.items()):
        if not want or name in want:
            run_model_case(name, cfg, "videoswintransformer")


if __name__ == "__main__":
    main(sys.argv[1:])
