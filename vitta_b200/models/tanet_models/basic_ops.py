"""Segment consensus -- mirror of the reference's ``models/tanet_models/basic_ops.py``."""
import torch


class Identity(torch.nn.Module):
    def forward(self, input):
        return input


class ConsensusModule(torch.nn.Module):
    """'avg': mean over the T axis of (N, T, K) frame scores, keepdim (reference :38-51,71-85; the
    hand-written autograd Function there is exactly the gradient of a mean).  'identity' / 'rnn': pass through."""

    def __init__(self, consensus_type, dim=1):
        super().__init__()
        self.consensus_type = consensus_type if consensus_type != 'rnn' else 'identity'
        self.dim = dim
        assert self.dim == 1

    def forward(self, input):
        if self.consensus_type == 'avg':
            return input.mean(dim=self.dim, keepdim=True)
        if self.consensus_type == 'identity':
            return input
        return None
