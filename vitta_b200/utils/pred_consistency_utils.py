"""Prediction-consistency loss -- mirror of the reference's ``utils/pred_consistency_utils.py``."""
from .. import ops


def compute_pred_consis(preds):
    """preds: (batch, n_views, n_class) logits.  loss = (1/V) sum_v || softmax(preds[:, v]) - mean_v softmax ||_1
    summed over batch and classes, the view-mean NOT detached (reference :15-31).  Forward and gradient come
    from one launch of the K10 kernel."""
    if preds.dim() != 3:
        raise ValueError("preds must be (batch, n_views, n_class)")
    return ops.PredConsisFn.apply(preds)
