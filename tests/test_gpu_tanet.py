"""GPU parity of the whole TANet adaptation step (fused sm_100a path behind the reference's driver API) against
the golden vectors recorded from the unmodified reference's ``tta_standard``."""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def _run_case(name, dev):
    import vitta_b200
    from vitta_b200 import synth
    from vitta_b200.corpus.basics import OnlineAdapter
    from vitta_b200.models.tanet_models.tanet import TSN
    from vitta_b200.utils import norm_stats_utils as nsu
    from vitta_b200.utils.opts import default_args
    vitta_b200.set_fp32_exact()
    nsu.reset_arenas()
    cfg = cases.TANET_CASES[name]
    g = cases.load_golden(name)
    src_m, src_v = cases.src_stats_from_golden(g)
    model = TSN(cfg["K"], cfg["T"], 'RGB', base_model='resnet50', consensus_type='avg', img_feature_dim=256,
                tam=True, non_local=False, partial_bn=False)
    sd0 = synth.synth_state_dict(model.state_dict(), seed=1)
    model.load_state_dict(sd0, strict=True)
    model.base_model.fc.p = 0.0
    model = model.to(dev)
    args = default_args(arch='tanet', clip_length=cfg["T"], batch_size=cfg["N"], n_augmented_views=cfg["M"],
                        if_pred_consistency=cfg["consis"], reg_type=cfg["reg_type"], lr=cfg["lr"],
                        num_classes=cfg["K"], input_size=cfg["res"], moving_avg=cfg["moving_avg"],
                        update_only_bn_affine=cfg.get("bn_affine", False), momentum_mvg=cfg.get("momentum_mvg", 0.1),
                        if_tta_standard=cfg.get("mode", "tta_online"), n_gradient_steps=cfg.get("gsteps", 1),
                        stat_reg=cfg.get("stat_reg", "mean_var"), before_norm=cfg.get("before_norm", False))
    bns = cfg.get("stat_reg") == "BNS"
    # the source statistics: our compute_statistics must reproduce the reference's (fused stats kernels, eval fwd)
    from vitta_b200.corpus.basics import compute_statistics
    clean = cases.case_inputs(cfg, "tanet", "clean", 2, 100)

    class DS(torch.utils.data.Dataset):
        def __init__(self, x):
            self.x = x

        def __len__(self):
            return self.x.shape[0]

        def __getitem__(self, i):
            return self.x[i], 0
    a2 = default_args(arch='tanet', clip_length=cfg["T"], batch_size=cfg["N"], num_classes=cfg["K"],
                      input_size=cfg["res"], stat_type='spatiotemp', result_dir=None,
                      before_norm=cfg.get("before_norm", False))
    a2.dataset_factory = lambda a, split, dataset_type: DS(torch.cat(clean, 0))
    om, ov = compute_statistics(model, a2)
    assert len(om) == len(src_m) == 53
    for i in range(len(om)):
        cases.assert_close(om[i], src_m[i], 2e-4, 2e-5 * float(np.abs(src_m[i]).max()) + 1e-6, "src_mean/%d" % i)
        cases.assert_close(ov[i], src_v[i], 2e-4, 2e-5 * float(np.abs(src_v[i]).max()) + 1e-6, "src_var/%d" % i)

    ad = OnlineAdapter(model, args, (src_m, src_v))
    assert len(ad.stat_reg_hooks) == int(g["n_hooks"]) == 47
    tta_in, eval_in = cases.tta_inputs(cfg, "tanet")
    gsteps = cfg.get("gsteps", 1)
    for s in range(cfg["steps"]):
        if cfg.get("mode") == "tta_standard" and s > 0:
            ad = OnlineAdapter(model, args, (src_m, src_v))   # tta_standard: fresh copy, optimiser and hooks per batch
        r = ad.adapt(tta_in[s].to(dev))
        # with n_gradient_steps > 1 the adapter reports the last gradient step of the batch
        pre = "step%d" % s if gsteps == 1 else "step%d.%d" % (s, gsteps - 1)
        # KLD divides by the adapted variance: the loss is ill-conditioned in the statistics (tests/cases.py
        # kld_loss_tolerance propagates the statistics' own tolerances through it); L1 / MSE are 1-Lipschitz-like
        kld_tol = cases.kld_loss_tolerance(g, pre, src_m, src_v) if cfg["reg_type"] == "kld" else None
        cases.assert_close(r["loss_reg"].cpu(), g[pre + "/loss_reg"], 1e-4,
                           1e-6 + (sum(kld_tol.values()) if kld_tol else 0.0), "loss_reg " + pre)
        if cfg["consis"]:
            cases.assert_close(r["loss_consis"].cpu(), g[pre + "/loss_consis"], 1e-3, 1e-7, "loss_consis")
        for h, hook in enumerate(ad.stat_reg_hooks):
            cases.assert_close(hook.r_feature.detach().cpu(), g["%s/r_feature/%d" % (pre, h)], 1e-4,
                               1e-6 + (kld_tol.get(h, 0.0) if kld_tol else 0.0), "r_feature %d" % h)
            k = "%s/ema_mean/%d" % (pre, h)
            if k in g:
                em, ev = g[k], g["%s/ema_var/%d" % (pre, h)]
                # per-channel means can be ~0, so the absolute floor is 3e-5 x the layer's activation
                # scale (|mean| + std over channels), not 1e-5 x max|mean|
                scale = float((np.abs(em) + np.sqrt(np.abs(ev))).max())
                got_m, got_v = (hook.mean, hook.var) if bns else (hook.ema_mean, hook.ema_var)
                cases.assert_close(got_m.cpu(), em, 1e-4, 3e-5 * scale + 1e-7, k)
                cases.assert_close(got_v.cpu(), ev, 1e-4, 1e-5 * float(np.abs(ev).max()) + 1e-7, "ema_var")
        ad.hooks_off()
        ev = ad.evaluate(eval_in[s].to(dev))
        want = g["step%d/eval_logits" % s]
        cases.assert_close(ev.cpu(), want, 2e-4, 2e-5 * float(np.abs(want).max()), "eval logits step %d" % s)
        ad.hooks_on()
    new_sd = ad.model.state_dict()
    names = [str(n) for n in g["delta_names"]]
    ref = g["delta_norm_sum"]
    for i, n in enumerate(names):
        d = (new_sd[n].detach().cpu() - sd0[n]).double()
        cases.assert_close(float(d.norm()), ref[i, 0], 1e-2, 1e-9, "delta norm " + n)
    for k in g.files:
        if k.startswith("delta/"):
            n = k[len("delta/"):]
            d = (new_sd[n].detach().cpu() - sd0[n]).reshape(-1)[:4096]
            scale = float(np.abs(g[k]).max()) + 1e-12
            # a delta is a difference of two fp32 weights: it is quantised to ulp(|w|) = 2^-23 |w|
            ulp = 1.2e-7 * float(sd0[n].abs().max())
            atol = 2e-3 * scale + ulp
            if cfg.get("bn_affine"):
                atol = cases.adam_delta_atol(g[k], cfg["lr"], cfg["steps"], atol)
            cases.assert_close(d, g[k], 5e-3, atol, k)


@pytest.mark.parametrize("name", ["tanet_t8_r64_consis_l1", "tanet_t8_r64_stats_mse", "tanet_t16_r224_stats_l1"])
def test_tanet_tta_vs_reference_golden(cuda_device, name):
    _run_case(name, cuda_device)


@pytest.mark.parametrize("name", ["tanet_t8_r64_stats_kld_avg", "tanet_t8_r64_consis_l1_bnaffine",
                                  "tanet_t8_r64_standard_l1", "tanet_t8_r64_bns_l1", "tanet_t8_r64_stats_l1_before_norm"])
def test_tanet_option_modes_vs_reference_golden(cuda_device, name):
    """SURVEY 8(f) rank 4 at model level: KLD + AverageMeterTensor statistics; --update_only_bn_affine (Adam);
    tta_standard mode (per-batch re-initialisation, momentum_mvg = 1, two gradient steps per batch); --stat_reg BNS;
    --before_norm (statistics of the norm inputs, through the generic forward-hook path)."""
    _run_case(name, cuda_device)


def test_state_dict_keys_match_reference_layout():
    """CPU-safe but cheap: also run on the GPU box so the driver sees the same module tree there."""
    from vitta_b200.models.tanet_models.tanet import TSN
    m = TSN(101, 8, 'RGB', base_model='resnet50', tam=True)
    tmpl = cases.tanet_state_template(101, 8)
    assert list(m.state_dict().keys()) == list(tmpl.keys())


def test_cuda_graph_replay_matches_eager(cuda_device):
    """args.cuda_graph: after 3 eager steps the step is captured and replayed; losses, logits and weights must follow the
    eager adapter exactly (same kernels in the same order; deterministic reductions)."""
    import vitta_b200
    from vitta_b200 import synth
    from vitta_b200.corpus.basics import OnlineAdapter
    from vitta_b200.models.tanet_models.tanet import TSN
    from vitta_b200.utils import norm_stats_utils as nsu
    from vitta_b200.utils.opts import default_args
    vitta_b200.set_fp32_exact()
    dev = cuda_device
    cfg = cases.TANET_CASES["tanet_t8_r64_stats_mse"]
    g = cases.load_golden("tanet_t8_r64_stats_mse")
    src_m, src_v = cases.src_stats_from_golden(g)
    outs = []
    for use_graph in (False, True):
        nsu.reset_arenas()
        model = TSN(cfg["K"], cfg["T"], 'RGB', base_model='resnet50', consensus_type='avg', img_feature_dim=256, tam=True,
                    non_local=False, partial_bn=False)
        model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=1), strict=True)
        model.base_model.fc.p = 0.0
        model = model.to(dev)
        args = default_args(arch='tanet', clip_length=cfg["T"], batch_size=cfg["N"], n_augmented_views=1,
                            if_pred_consistency=False, reg_type="l1_loss", lr=1e-3, num_classes=cfg["K"],
                            input_size=cfg["res"], moving_avg=True, cuda_graph=use_graph)
        ad = OnlineAdapter(model, args, (src_m, src_v))
        rec = []
        for s in range(7):
            x = synth.tanet_loader_tensor(synth.synth_video(cfg["N"], 1, cfg["T"], cfg["res"], seed=500 + s, tag="tta")).to(dev)
            r = ad.adapt(x)
            rec.append((float(r["loss_reg"]), r["output"].clone().cpu()))
        if use_graph:
            assert ad._graph is not None, "the step was never captured"
        ad.hooks_off()
        ev = ad.evaluate(x).clone().cpu()
        if use_graph:      # the evaluation forward: eager pass, capture + first replay, replay -- all the same logits
            for _ in range(2):
                ev2 = ad.evaluate(x).clone().cpu()
                torch.testing.assert_close(ev2, ev, rtol=1e-6, atol=1e-8)
            assert getattr(ad, "_eval_graph", None) is not None, "the evaluation forward was never captured"
        outs.append((rec, ev, ad.model.new_fc.weight.detach().cpu().clone(),
                     ad.model.base_model.layer3[2].net.conv2.weight.detach().cpu().clone()))
    (re, eve, w1e, w2e), (rg, evg, w1g, w2g) = outs
    for s, ((le, oe), (lg, og)) in enumerate(zip(re, rg)):
        assert abs(le - lg) <= 1e-6 * abs(le) + 1e-9, (s, le, lg)
        torch.testing.assert_close(og, oe, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(evg, eve, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(w2g, w2e, rtol=1e-5, atol=1e-9)


def test_tanet_golden_under_both_operand_splits(cuda_device):
    """The library default is the fp16 hi/lo split (f16x3); the tf32 split stays selectable.  The small reference golden
    must hold under BOTH (the other goldens of this file run under the default)."""
    from vitta_b200 import ops
    before = ops.gemm_precision()
    try:
        for prec in ("tf32x3", "f16x3"):
            ops.set_gemm_precision(prec)
            _run_case("tanet_t8_r64_stats_mse", cuda_device)
    finally:
        ops.set_gemm_precision(before)


def test_bn_folded_inference_forward_matches_layerwise_forward(cuda_device, monkeypatch):
    """The evaluation forward folds every eval-mode BatchNorm into its convolution (K6 epilogue: + bias, + shortcut, ReLU;
    two multi-tensor launches refresh all folded operands after a weight update).  It must agree with the layer-by-layer
    forward (VITTA_INFER_FOLD=0) to fp32 rounding, before and after an adaptation step changed the weights."""
    import vitta_b200
    from vitta_b200 import synth
    from vitta_b200.corpus.basics import OnlineAdapter
    from vitta_b200.models.tanet_models.tanet import TSN
    from vitta_b200.utils import norm_stats_utils as nsu
    from vitta_b200.utils.opts import default_args
    from oracle import vitta_oracle as O
    vitta_b200.set_fp32_exact()
    nsu.reset_arenas()
    K, T, N, res = 11, 8, 2, 64
    model = TSN(K, T, 'RGB', base_model='resnet50', consensus_type='avg', img_feature_dim=256, tam=True, non_local=False,
                partial_bn=False)
    sd = synth.synth_state_dict(model.state_dict(), seed=1)
    model.load_state_dict(sd)
    model.base_model.fc.p = 0.0
    model = model.to(cuda_device)
    clean = synth.tanet_loader_tensor(synth.synth_video(N, 1, T, res, seed=100, gauss_sigma=0.0, tag="clean"))
    src_m, src_v = O.collect_source_stats(sd, "tanet", T, [clean.view(N, T, 3, res, res)])
    args = default_args(arch='tanet', clip_length=T, batch_size=N, n_augmented_views=1, if_pred_consistency=False,
                        lr=1e-3, num_classes=K, input_size=res)
    ad = OnlineAdapter(model, args, (src_m, src_v))
    x = synth.tanet_loader_tensor(synth.synth_video(N, 1, T, res, seed=200, tag="tta")).to(cuda_device)
    for step in range(2):
        monkeypatch.setenv("VITTA_INFER_FOLD", "1")
        a = ad.evaluate(x)
        assert getattr(ad.model.base_model, "_folds", None) is not None          # the folded path really ran
        monkeypatch.setenv("VITTA_INFER_FOLD", "0")
        b = ad.evaluate(x)
        cases.assert_close(a.cpu(), b.cpu().numpy(), 1e-4, 1e-5 * float(b.abs().max()), "folded vs layerwise logits, step %d" % step)
        ad.hooks_on()
        ad.adapt(x)                                                               # weights move: folds must refresh
