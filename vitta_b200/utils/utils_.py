"""Meters and accuracy -- mirror of the small host-side helpers of the reference's ``utils/utils_.py``.

The alignment hooks do NOT use these classes on the hot path (their EMA lives in the K2 finalize kernel);
they are kept because ``corpus.basics`` and user code construct them by name."""
import logging
import os
import sys
import time

import torch


class AverageMeter(object):
    """Running weighted mean of python/torch scalars (reference utils/utils_.py:171-187)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class AverageMeterTensor(object):
    """Weighted running mean whose history is detached (reference :190-202)."""

    def __init__(self, device=None):
        self.device = device
        self.reset()

    def reset(self):
        z = torch.zeros((), dtype=torch.float32, device=self.device)
        self.val, self.avg, self.sum, self.count = z, z.clone(), z.clone(), 0

    def update(self, val, n=1):
        self.val = val
        self.sum = self.sum.detach().to(val.device) + val * n
        self.count += n
        self.avg = self.sum / self.count


class MovingAverageTensor(object):
    """EMA that starts from the scalar 0 (no bias correction) with detached history (reference :204-211)."""

    def __init__(self, momentum=0.1, device=None):
        self.momentum = momentum
        self.device = device
        self.reset()

    def reset(self):
        self.avg = torch.zeros((), dtype=torch.float32, device=self.device)

    def update(self, val):
        self.avg = self.momentum * val + (1.0 - self.momentum) * self.avg.detach().to(val.device)


def accuracy(output, target, topk=(1,)):
    """precision@k in percent (reference :224-237)."""
    kmax = max(topk)
    n = target.size(0)
    idx = output.topk(kmax, dim=1, largest=True, sorted=True).indices        # (n, kmax)
    hit = idx.eq(target.view(-1, 1))
    return [hit[:, :k].any(dim=1).float().sum().mul_(100.0 / n) for k in topk]


def make_dir(path):
    os.makedirs(path, exist_ok=True)


def path_logger(result_dir, log_time):
    """File + stream logger (reference :92-110)."""
    make_dir(result_dir)
    logger = logging.getLogger("vitta_b200.%s.%s" % (result_dir, log_time))
    logger.setLevel(logging.DEBUG)
    logger.propagate = False
    fmt = logging.Formatter("%(asctime)s %(levelname)s %(message)s")
    fh = logging.FileHandler(os.path.join(result_dir, str(log_time)))
    fh.setFormatter(fmt)
    sh = logging.StreamHandler(sys.stderr)
    sh.setFormatter(fmt)
    logger.addHandler(fh)
    logger.addHandler(sh)
    return logger


def get_writer_to_all_result(args, custom_path=None):
    """The per-run results file of the entry scripts (reference :252-267): ``<result_dir>/<time>_all_result`` (or
    ``<custom_path>/<baseline>_<time>_all_result``) opened 'w+', a header with one ``name value`` line per public attribute
    of ``args``, two separator lines and two blank lines; the caller then appends one line of accuracies per corruption."""
    log_time = time.strftime("%Y%m%d_%H%M%S")
    if custom_path is None:
        make_dir(args.result_dir)       # eval() normally created it already; the reference relies on that
        f_write = open(os.path.join(args.result_dir, f'{log_time}_all_result'), 'w+')
    else:
        make_dir(custom_path)
        f_write = open(os.path.join(custom_path, f'{args.baseline}_{log_time}_all_result'), 'w+')
    for arg in dir(args):
        if arg[0] != '_':
            f_write.write(f'{arg} {getattr(args, arg)}\n')
    f_write.write('#############################\n')
    f_write.write('#############################\n')
    f_write.write('\n')
    f_write.write('\n')
    return f_write
