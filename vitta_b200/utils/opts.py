"""Option surface -- mirror of the reference's ``utils/opts.py`` (same names and defaults, table-driven here).

Preserved pitfalls: ``type=bool`` flags parse any non-empty string as True; options without a
``type`` (``--lr``, ``--n_epoch_adapat``, ``--patch_size``, ``--window_size``, ``--chosen_blocks``, ``--stat_type``,
``--tta_view_sample_style_list``) are only usefully set from Python, which is what the entry scripts do."""
import argparse

input_mean = [0.485, 0.456, 0.406]
input_std = [0.229, 0.224, 0.225]
img_norm_cfg = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_bgr=False)

_STR, _INT, _FLOAT, _BOOL, _FLAG, _RAW = "str", "int", "float", "bool", "flag", "raw"
_TYPES = {_STR: str, _INT: int, _FLOAT: float, _BOOL: bool}

# (flags, kind, default, extra)
_OPTIONS = [
    # data
    (("--dataset",), _STR, "ucf101", dict(choices=["ucf101", "somethingv2", "kinetics"])),
    (("--modality",), _STR, "RGB", {}),
    (("--root_path",), _STR, "None", {}),
    (("--video_data_dir",), _STR, "", {}),
    (("--vid_format",), _STR, "", {}),
    (("--datatype",), _STR, "vid", dict(choices=["vid", "frame"])),
    (("--spatiotemp_mean_clean_file",), _STR, "", {}),
    (("--spatiotemp_var_clean_file",), _STR, "", {}),
    (("--val_vid_list",), _STR, "{}.txt", {}),
    (("--result_dir",), _STR, "results/{}_{}/tta_{}", {}),
    # model
    (("--arch",), _STR, "tanet", dict(choices=["tanet", "videoswintransformer"])),
    (("--model_path",), _STR, "", {}),
    (("--img_feature_dim",), _INT, 256, {}),
    (("--partial_bn",), _FLAG, False, {}),
    # video swin
    (("--num_clips",), _INT, 1, {}),
    (("--frame_uniform",), _BOOL, True, {}),
    (("--frame_interval",), _INT, 2, {}),
    (("--flip_ratio",), _INT, 0, {}),
    (("--img_norm_cfg",), _RAW, img_norm_cfg, {}),
    (("--patch_size",), _RAW, (2, 4, 4), {}),
    (("--window_size",), _RAW, (8, 7, 7), {}),
    (("--drop_path_rate",), _RAW, 0.2, {}),
    # runtime
    (("--gpus",), _INT, None, dict(nargs="+")),
    (("-j", "--workers"), _INT, 8, {}),
    (("--norm",), _FLAG, False, {}),
    (("--debug",), _FLAG, False, {}),
    (("--verbose",), _BOOL, True, {}),
    (("--print-freq", "-p"), _INT, 20, {}),
    # learning
    (("--tta",), _BOOL, True, {}),
    (("--use_src_stat_in_reg",), _BOOL, True, {}),
    (("--fix_BNS",), _BOOL, True, {}),
    (("--running_manner",), _BOOL, True, {}),
    (("--momentum_bns",), _FLOAT, 0.1, {}),
    (("--update_only_bn_affine",), _FLAG, False, {}),
    (("--compute_stat",), _FLAG, False, {}),
    (("--momentum_mvg",), _FLOAT, 0.1, {}),
    (("--stat_reg",), _STR, "mean_var", {}),
    (("--if_tta_standard",), _STR, "tta_online", {}),
    (("--loss_type",), _STR, "nll", dict(choices=["nll"])),
    (("--if_sample_tta_aug_views",), _BOOL, True, {}),
    (("--if_spatial_rand_cropping",), _BOOL, True, {}),
    (("--if_pred_consistency",), _BOOL, True, {}),
    (("--lambda_pred_consis",), _FLOAT, 0.1, {}),
    (("--lambda_feature_reg",), _INT, 1, {}),
    (("--n_augmented_views",), _INT, 2, {}),
    (("--tta_view_sample_style_list",), _RAW, ["uniform_equidist"], {}),
    (("--stat_type",), _RAW, ["spatiotemp"], {}),
    (("--before_norm",), _FLAG, False, {}),
    (("--reduce_dim",), _BOOL, True, {}),
    (("--reg_type",), _STR, "l1_loss", {}),
    (("--chosen_blocks",), _RAW, ["layer3", "layer4"], {}),
    (("--moving_avg",), _BOOL, True, {}),
    (("--n_gradient_steps",), _INT, 1, {}),
    (("--full_res",), _FLAG, False, {}),
    (("--input_size",), _INT, 224, {}),
    (("--scale_size",), _INT, 256, {}),
    (("--batch_size",), _INT, 1, {}),
    (("--clip_length",), _INT, 16, {}),
    (("--sample_style",), _STR, "uniform-1", {}),
    (("--test_crops",), _INT, 1, {}),
    (("--use_pretrained",), _FLAG, False, {}),
    (("--input_mean",), _RAW, input_mean, {}),
    (("--input_std",), _RAW, input_std, {}),
    (("--lr",), _RAW, 0.00005, {}),
    (("--n_epoch_adapat",), _RAW, 1, {}),
    (("--momentum",), _FLOAT, 0.9, {}),
    (("--weight-decay", "--wd"), _FLOAT, 5e-4, {}),
]


def _build_parser():
    p = argparse.ArgumentParser(description="ViTTA")
    for flags, kind, default, extra in _OPTIONS:
        if kind == _FLAG:
            p.add_argument(*flags, action="store_true")
        elif kind == _RAW:
            p.add_argument(*flags, default=default)
        else:
            p.add_argument(*flags, type=_TYPES[kind], default=default, **extra)
    return p


parser = _build_parser()


def get_opts(argv=None):
    args = parser.parse_args(argv)
    args.evaluate_baselines = not args.tta
    args.baseline = 'source'
    return args


def default_args(**overrides):
    """Programmatic construction (tests, bench): defaults of the parser + keyword overrides."""
    args = get_opts([])
    args.workers = 0
    args.verbose = False
    args.num_classes = 101
    for k, v in overrides.items():
        setattr(args, k, v)
    return args
