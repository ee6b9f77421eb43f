"""ddm/loss.py:292-382 — MSE_Loss / MAE_Loss (non-mask branches), the per-sample reductions DDM uses.  In the fused
training path these are evaluated inside adm_ddm_loss (K2); the classes exist so that ``loss_main`` configs resolve."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MSE_Loss(nn.Module):
    def __init__(self, thresh_min=0, thresh_max=1, mask=False, with_sigmoid=False):
        super().__init__()
        if mask:
            raise NotImplementedError("masked MSE_Loss (depth estimation) is outside the DDM hot path")
        self.with_sigmoid = with_sigmoid

    def forward(self, pred, gt, reduce_dims=[1, 2, 3], mask=None, reduction="mean"):
        if self.with_sigmoid:
            pred, gt = torch.sigmoid(pred), torch.sigmoid(gt)
        loss = F.mse_loss(pred, gt, reduction="none")
        if reduction == "mean":
            return loss.mean(dim=reduce_dims)
        if reduction == "sum":
            return loss.sum(dim=reduce_dims)
        if reduction == "none":
            return loss
        raise NotImplementedError("")


class MAE_Loss(nn.Module):
    def __init__(self, thresh_min=0, thresh_max=1, mask=False, with_sigmoid=False):
        super().__init__()
        if mask:
            raise NotImplementedError("masked MAE_Loss is outside the DDM hot path")

    def forward(self, pred, gt, reduce_dims=[1, 2, 3], mask=None, reduction="mean"):
        loss = F.l1_loss(pred, gt, reduction="none")
        if reduction == "mean":
            return loss.mean(dim=reduce_dims)
        if reduction == "sum":
            return loss.sum(dim=reduce_dims)
        return loss
