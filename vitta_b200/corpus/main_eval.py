"""Dispatcher -- mirror of the hot-path branches of the reference's ``corpus/main_eval.py::eval`` (:30-232):
build / load the model, wrap it so module names carry the ``module.`` prefix the hook attachment matches on
(reference :61-65; ``chosen_blocks`` of the Swin script, tta_swin_ucf101.py:40), then dispatch to
``compute_statistics`` (``--compute_stat mean_var``), ``tta_standard`` (``--tta True``) or source-only ``validate``.

Out of scope: the competing TTA baselines (``norm``, ``tent``, ``shot``, ``dua``, ``t3a``) and the
cosine-similarity statistics -- they raise NotImplementedError.
"""
import os.path as osp
import time

import torch
import torch.backends.cudnn as cudnn

from .. import set_fp32_exact
from ..utils.utils_ import make_dir, path_logger
from .basics import (_loader, compute_statistics, get_dataset_tanet, get_dataset_videoswin, get_model, tta_standard,
                     validate)

NUM_CLASSES = {'ucf101': 101, 'hmdb51': 51, 'kinetics': 400, 'somethingv2': 174, 'kth': 6, 'u2h': 12, 'h2u': 12}


def load_checkpoint_into(model, args, logger=None):
    """reference :55-65: ``state_dict`` key, optionally ``module.``-prefixed; returns the DataParallel-wrapped model."""
    checkpoint = torch.load(args.model_path, map_location='cpu')
    if logger is not None:
        logger.debug(f'Loading {args.model_path}')
    if args.arch == 'tanet' and 'epoch' in checkpoint:
        print("model epoch {} best prec@1: {}".format(checkpoint['epoch'], checkpoint.get('best_prec1')))
    sd = checkpoint['state_dict']
    if 'module.' in list(sd.keys())[0]:
        model = torch.nn.DataParallel(model, device_ids=args.gpus).cuda()
        model.load_state_dict(sd)
    else:
        model.load_state_dict(sd)
        model = torch.nn.DataParallel(model, device_ids=args.gpus).cuda()
    return model


def _maybe_init_distributed(args):
    """torchrun launch (WORLD_SIZE > 1): one process per GPU over NCCL; ``tta_standard`` then shards the videos of every
    loader batch over the ranks and the step's two collectives keep statistics and weights identical everywhere
    (DESIGN.md section 7).  The reference itself is single-process; without torchrun nothing here runs."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or getattr(args, 'process_group', None) is not None:
        return
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.process_group = dist.group.WORLD
    args.gpus = [local]


def resolve_dataset_factory(args):
    """``VITTA_DATASET_FACTORY=package.module:callable`` (or ``args.dataset_factory`` set by the caller) plugs a real
    loader -- ``callable(args, split, dataset_type) -> torch Dataset`` in the reference's loader layout, e.g.
    ``vitta_b200.corpus.views:DecodedVideoDataset.factory`` -- into the entry scripts."""
    import importlib
    import os
    spec = os.environ.get("VITTA_DATASET_FACTORY")
    if getattr(args, 'dataset_factory', None) is None and spec:
        mod, _, attr = spec.partition(":")
        obj = importlib.import_module(mod)
        for part in attr.split("."):
            obj = getattr(obj, part)
        args.dataset_factory = obj
    return getattr(args, 'dataset_factory', None)


def check_inputs_or_synthetic(args, model_given):
    """The run needs a checkpoint, a dataset and (for mean_var alignment) source statistics.  Anything missing is an
    error unless the caller explicitly asks for the synthetic stand-ins (``args.synthetic`` / ``VITTA_SYNTHETIC=1``:
    seeded random weights, synthetic videos, statistics fabricated from a synthetic clean set) -- a result file
    written from stand-ins must never look like a measured accuracy."""
    import os
    import sys
    missing = []
    if not model_given and not getattr(args, 'model_path', None):
        missing.append("model_path (checkpoint)")
    if resolve_dataset_factory(args) is None:
        missing.append("dataset (args.dataset_factory / VITTA_DATASET_FACTORY)")
    if (args.tta and args.stat_reg == 'mean_var' and args.compute_stat in (False, 'False')
            and getattr(args, 'source_stats', None) is None and not getattr(args, 'spatiotemp_mean_clean_file', None)):
        missing.append("source statistics (spatiotemp_mean/var_clean_file)")
    if not missing:
        return []
    if not (getattr(args, 'synthetic', False) or os.environ.get("VITTA_SYNTHETIC") == "1"):
        raise RuntimeError("missing inputs: %s.  Provide them, or set VITTA_SYNTHETIC=1 (args.synthetic=True) to run on "
                           "seeded random weights / synthetic videos / fabricated statistics on purpose."
                           % "; ".join(missing))
    sys.stderr.write("vitta_b200: SYNTHETIC RUN -- stand-ins used for: %s. Accuracies are meaningless.\n"
                     % "; ".join(missing))
    return missing


def eval(args=None, model=None):
    log_time = time.strftime("%Y%m%d_%H%M%S")
    make_dir(args.result_dir)
    logger = path_logger(args.result_dir, log_time)
    if args.verbose:
        for arg in dir(args):
            if arg[0] != '_':
                logger.debug(f'{arg} {getattr(args, arg)}')
    num_classes = NUM_CLASSES[args.dataset]
    args.num_classes = num_classes
    if not torch.cuda.is_available():
        raise RuntimeError("vitta_b200 needs a CUDA device (sm_100a); there is no CPU path")
    set_fp32_exact()
    args.synthetic_stand_ins = check_inputs_or_synthetic(args, model is not None)
    if args.synthetic_stand_ins:
        logger.debug('SYNTHETIC RUN, stand-ins for: ' + '; '.join(args.synthetic_stand_ins))
    _maybe_init_distributed(args)
    if model is None:
        if not getattr(args, 'model_path', None):
            # synthetic stand-in for the checkpoint: the SAME seeded initialisation in every process and run (under
            # torchrun all ranks must start from identical weights, as they would after loading one checkpoint)
            torch.manual_seed(int(getattr(args, 'synthetic_seed', 0)))
        model = get_model(args, num_classes, logger)
        if getattr(args, 'model_path', None):
            model = load_checkpoint_into(model, args, logger)
        else:   # no checkpoint (offline benchmarking): seeded random initialisation
            model = torch.nn.DataParallel(model, device_ids=args.gpus).cuda()
    args.crop_size = args.input_size
    cudnn.benchmark = True
    if args.loss_type == 'nll':
        criterion = torch.nn.CrossEntropyLoss().cuda()
    else:
        raise ValueError("Unknown loss type")
    epoch_result_list = None
    if args.tta:
        if args.compute_stat == 'mean_var':
            compute_statistics(model, args=args, logger=logger, log_time=log_time)
        elif args.compute_stat == 'cossim':
            raise NotImplementedError("relation-map statistics are outside the ViTTA hot path (SURVEY.md section 2)")
        elif args.compute_stat is False or args.compute_stat == 'False':
            if (args.stat_reg == 'mean_var' and getattr(args, 'source_stats', None) is None
                    and not getattr(args, 'spatiotemp_mean_clean_file', None)):
                # offline run without the reference's statistics files: fabricate them with the compute_stats/ flow on a
                # clean synthetic set (same producer, same list format)
                logger.debug('no source statistics given: computing them on the synthetic clean set')
                import copy
                a2 = copy.copy(args)
                a2.stat_type = 'spatiotemp'
                a2.result_dir = None
                a2.synthetic_seed = getattr(args, 'synthetic_seed', 0) + 1000
                args.source_stats = compute_statistics(model, args=a2, logger=logger, log_time=log_time)
            if args.if_tta_standard:
                epoch_result_list = tta_standard(model, criterion, args=args, logger=logger, writer=None)
                model = None
            else:
                raise NotImplementedError("test_time_adapt is unreachable from the shipped scripts (if_tta_standard is "
                                          "always truthy, utils/opts.py:81)")
    elif args.evaluate_baselines:
        if args.baseline != 'source':
            raise NotImplementedError(f"baseline {args.baseline!r}: only source-only evaluation is on the hot path")
        make = get_dataset_tanet if args.arch == 'tanet' else get_dataset_videoswin
        val_loader = _loader(make(args, split='val', dataset_type='eval'), args)
        top1_acc = validate(val_loader, model, criterion, 0, epoch=0, args=args, logger=logger)
        epoch_result_list = [top1_acc]
    return epoch_result_list, model
