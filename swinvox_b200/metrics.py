"""Threshold + IoU / F-score of the evaluation loop (reference core/test.py:141-164), batched on the device:
one kernel produces the integer counters {I, U, TP, FP, FN} per object and threshold; the float formulas (with the
reference's epsilons and its union == 0 convention) are applied to those counters on the host."""
import torch

from . import engine as E
from .models import _base


class VoxelMetrics:
    def __init__(self, thresholds=(0.2, 0.3, 0.4, 0.5)):
        self.thresholds = [float(t) for t in thresholds]
        self._plans = {}

    def counts_and_bce(self, logits, gt):
        """-> (counts int32 [B,T,5], bce fp64 [B]): bce[b] = mean over voxels of BCEWithLogits(logits[b], gt[b]), the
        per-sample EDLoss / RLoss of core/test.py:133-139 before its factor 10 (batch size 1 there)"""
        counts = self.counts(logits, gt, want_bce=True)
        key = self._last_key
        return counts, self._plans[key][2].to(torch.float64) / (1048576.0 * logits[0].numel())

    def counts(self, logits, gt, want_bce=False):
        """logits, gt: [B,32,32,32] (or [B,P]) fp32 on the GPU -> int32 [B, T, 5] = I, U, TP, FP, FN (device)"""
        _base.require_device(logits)
        B = logits.shape[0]
        P = logits[0].numel()
        if not (logits.is_contiguous() and gt.is_contiguous()):   # reshape would silently bind the plan to a copy
            raise ValueError("VoxelMetrics needs contiguous logits / ground truth")
        if logits.dtype != torch.float32 or gt.dtype != torch.float32:
            raise TypeError("VoxelMetrics needs float32 logits / ground truth")
        key = (B, P, str(logits.device), logits.data_ptr(), gt.data_ptr(), bool(want_bce))
        self._last_key = key
        if key not in self._plans:
            plan = E.Plan(logits.device)
            th = torch.tensor(self.thresholds, dtype=torch.float32, device=logits.device)
            counts = plan.zeros(B, len(self.thresholds), 5, dtype=torch.int32)
            lg, g = logits.view(B, P), gt.view(B, P)
            bce = plan.zeros(B, dtype=torch.int64) if want_bce else None
            plan.voxel_metrics(lg, g, th, counts, B, P, bce=bce)
            if len(self._plans) > 8:
                self._plans.clear()
            self._plans[key] = (plan, counts, bce)
        plan, counts, _ = self._plans[key]
        plan.run()
        return counts

    @staticmethod
    def scores(counts):
        """counts [B,T,5] -> (iou [B,T], f1 [B,T]) with the reference's conventions (core/test.py:150-164)"""
        c = counts.to(torch.float32)
        inter, union, tp, fp, fn = c.unbind(-1)
        iou = torch.where(union > 0, inter / union.clamp_min(1), torch.where(inter == 0, torch.ones_like(inter),
                                                                             torch.zeros_like(inter)))
        prec, rec = tp / (tp + fp + 1e-8), tp / (tp + fn + 1e-8)
        return iou, 2 * prec * rec / (prec + rec + 1e-8)
