"""Segmentation scores of the validation loop (reference utils.py:357-412 `runningScore`, used at
model.py:555-572): confusion matrix accumulated ON THE DEVICE for CUDA label maps (sscg_confusion), with numpy
for CPU inputs exactly like the reference; `get_scores()` follows the reference line by line, including its
per-dataset class exclusions (VOC ignores class 0, Cityscapes the last class)."""
import numpy as np
import torch

from . import _lib as L
from .kernels import _ptr, _stream


class RunningScore:
    def __init__(self, n_classes, dataset):
        self.n_classes = n_classes
        self.dataset = dataset
        self.confusion_matrix = np.zeros((n_classes, n_classes))        # utils.py:360
        self._dev_hist = None                                           # int64 [C, C] on the device, folded in lazily

    def _fast_hist(self, label_true, label_pred, n_class):              # utils.py:363-369
        mask = (label_true >= 0) & (label_true < n_class)
        hist = np.bincount(n_class * label_true[mask].astype(int) + label_pred[mask],
                           minlength=n_class ** 2).reshape(n_class, n_class)
        return hist

    def update(self, label_trues, label_preds):                         # utils.py:371-375
        if torch.is_tensor(label_trues) and label_trues.is_cuda:
            lt = label_trues.reshape(-1).to(torch.int64).contiguous()
            lp = label_preds.reshape(-1).to(torch.int64).contiguous()
            assert lt.numel() == lp.numel() and lp.is_cuda
            if self._dev_hist is None:
                self._dev_hist = torch.zeros(self.n_classes, self.n_classes, dtype=torch.int64, device=lt.device)
            L.check(L.lib().sscg_confusion(_ptr(lt), _ptr(lp), lt.numel(), self.n_classes, _ptr(self._dev_hist), _stream()),
                    "sscg_confusion")
            return
        if torch.is_tensor(label_trues):
            label_trues, label_preds = label_trues.numpy(), label_preds.numpy()
        for lt, lp in zip(label_trues, label_preds):
            self.confusion_matrix += self._fast_hist(lt.flatten(), lp.flatten(), self.n_classes)

    def _fold(self):
        if self._dev_hist is not None:
            self.confusion_matrix += self._dev_hist.cpu().numpy().astype(np.float64)
            self._dev_hist.zero_()

    def get_scores(self):                                               # utils.py:377-408
        self._fold()
        hist = self.confusion_matrix
        n = self.n_classes
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(hist).sum() / hist.sum()
            acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
            if self.dataset == "voc2012":
                h = hist[1:n, 1:n]
            elif self.dataset == "cityscapes":
                h = hist[0:n - 1, 0:n - 1]
            elif self.dataset == "acdc":
                h = hist
            else:
                raise ValueError("dataset must be 'voc2012', 'cityscapes' or 'acdc'")
            iu = np.diag(h) / (h.sum(axis=1) + h.sum(axis=0) - np.diag(h))
        mean_iu = np.nanmean(iu)
        cls_iu = dict(zip(range(n if self.dataset == "acdc" else n - 1), iu))
        return ({"Overall Acc: \t": acc, "Mean Acc : \t": acc_cls, "Mean IoU : \t": mean_iu}, cls_iu)

    def reset(self):                                                    # utils.py:410-411
        self.confusion_matrix = np.zeros((self.n_classes, self.n_classes))
        if self._dev_hist is not None:
            self._dev_hist.zero_()
