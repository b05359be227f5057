"""Drop-in replacements for the reference's nn.Modules on the hot path.

Same class names, constructor arguments, parameter names (state_dict-compatible with the
reference's `synthesizer_*.pth` checkpoints, SoP++/main.py:616-617) and call signatures as
models/synthesizer_net.py and models/criterion.py; the arithmetic runs in the CUDA library.
"""
from __future__ import annotations


import torch
import torch.nn as nn

from . import ops


def activate(x, activation):
    """models/__init__.py:16-28.  sigmoid / identity only: the other activations are off the hot path."""
    if activation == 'sigmoid':
        return torch.sigmoid(x)
    elif activation == 'no':
        return x
    elif activation in ('softmax', 'relu', 'tanh'):
        return {'softmax': lambda t: torch.softmax(t, dim=1), 'relu': torch.relu, 'tanh': torch.tanh}[activation](x)
    else:
        raise Exception('Unkown activation!')


class InnerProd(nn.Module):
    """models/synthesizer_net.py:6-38.  z[b,0,h,w] = sum_k feat_img[b,k] * scale[k] * feat_sound[b,k,h,w] + bias."""

    def __init__(self, fc_dim):
        super(InnerProd, self).__init__()
        self.scale = nn.Parameter(torch.ones(fc_dim))
        self.bias = nn.Parameter(torch.zeros(1))

    def forward(self, feat_img, feat_sound):
        B, C = feat_sound.size(0), feat_sound.size(1)
        z = ops.synth_masks(feat_sound, feat_img.reshape(B, 1, C), self.scale, self.bias, activation='no')
        return z[0].unsqueeze(1)

    # B200-first extensions: all N sources of the mixture in ONE pass over feat_sound -------------------
    def forward_multi(self, ctx_feats, feat_sound, activation='sigmoid'):
        """ctx_feats [B,N,K] -> list of N masks [B,1,H,W] = activate(self(ctx_feats[:, n], feat_sound), activation)
        (replaces the per-source loop at SoP++/main.py:203-206)."""
        z = ops.synth_masks(feat_sound, ctx_feats, self.scale, self.bias, activation=activation)
        return [z[n].unsqueeze(1) for n in range(z.shape[0])]

    def forward_loss(self, ctx_feats, feat_sound, gt_masks, weight, loss='bce', activation='sigmoid'):
        """Fused masks + criterion (SoP++/main.py:203-209): returns (err, list of N detached masks [B,1,H,W])."""
        err, z = ops.synth_masks_loss(feat_sound, ctx_feats, self.scale, self.bias, gt_masks, weight,
                                      loss=loss, activation=activation)
        return err, [z[n].unsqueeze(1) for n in range(z.shape[0])]

    def forward_pit(self, ctx_feats, feat_sound, gt_masks, weight, loss='bce', activation='sigmoid'):
        """Fused audio-only step (SoP++/main.py:227-238: synthesizer loop + sigmoid + PitWrapper + reorder_tensor): returns
        (err [B], perms, ordered masks as a list of N [B,1,H,W]); one pass over feat_sound forward, one backward."""
        err, perms, ordered, _ = ops.synth_masks_pit_loss(feat_sound, ctx_feats, self.scale, self.bias, gt_masks, weight,
                                                          loss=loss, activation=activation)
        return err, perms, [ordered[n].unsqueeze(1) for n in range(ordered.shape[0])]

    def forward_nosum(self, feat_img, feat_sound):
        return ops.synth_nosum(feat_img, feat_sound, self.scale, self.bias)

    # inference purposes
    def forward_pixelwise(self, feats_img, feat_sound):
        return ops.synth_pixelwise(feats_img, feat_sound, self.scale, self.bias)


class Bias(nn.Module):
    """models/synthesizer_net.py:41-70: InnerProd without the per-channel scale."""

    def __init__(self):
        super(Bias, self).__init__()
        self.bias = nn.Parameter(torch.zeros(1))

    def forward(self, feat_img, feat_sound):
        B, C = feat_sound.size(0), feat_sound.size(1)
        z = ops.synth_masks(feat_sound, feat_img.reshape(B, 1, C), None, self.bias, activation='no')
        return z[0].unsqueeze(1)

    def forward_multi(self, ctx_feats, feat_sound, activation='sigmoid'):
        z = ops.synth_masks(feat_sound, ctx_feats, None, self.bias, activation=activation)
        return [z[n].unsqueeze(1) for n in range(z.shape[0])]

    def forward_nosum(self, feat_img, feat_sound):
        return ops.synth_nosum(feat_img, feat_sound, None, self.bias)

    # inference purposes
    def forward_pixelwise(self, feats_img, feat_sound):
        return ops.synth_pixelwise(feats_img, feat_sound, None, self.bias)


class BaseLoss(nn.Module):
    """models/criterion.py:6-25."""
    kind = None

    def __init__(self):
        super(BaseLoss, self).__init__()

    def forward(self, preds, targets, weight=None):
        if isinstance(preds, list):
            return ops.list_loss(self.kind, preds, targets, weight)
        elif isinstance(preds, torch.Tensor):
            return ops.list_loss(self.kind, [preds], [targets], weight)
        raise TypeError("preds must be a list of tensors or a tensor")


class L1Loss(BaseLoss):
    """models/criterion.py:28-33."""
    kind = "l1"


class L2Loss(BaseLoss):
    """models/criterion.py:36-41."""
    kind = "l2"


class BCELoss(BaseLoss):
    """models/criterion.py:44-49."""
    kind = "bce"


class PitWrapper(nn.Module):
    """models/criterion.py:74-231 with base_loss = F.binary_cross_entropy (models/__init__.py:130-131).

    forward(preds [B,...,S], targets, weights) -> (loss [B], perms): the S x S loss matrix of every sample is
    computed by one kernel over the whole batch (the reference loops over samples and materialises [H,W,S,S]
    repeats), the S! search by a second kernel on the device (no D2H, no stream synchronisation: `perms` is a
    device-backed sequence that only copies when Python reads its values), and the gradient flows back to `preds`
    through one streaming kernel.  `base_loss` may be the string 'bce' / 'l1' / 'l2' or F.binary_cross_entropy (as the
    reference passes it)."""

    def __init__(self, base_loss="bce"):
        super(PitWrapper, self).__init__()
        self.kind = base_loss if isinstance(base_loss, str) else "bce"

    def reorder_tensor(self, tensor, p):
        """criterion.py:180-200; `p` is what forward() returned (device-backed) or any list of tuples."""
        return ops.pit_reorder(tensor, p)

    def forward(self, preds, targets, weights=None):
        if weights is None:
            weights = torch.ones_like(preds)
        return ops.pit_loss(self.kind, preds, targets, weights)
