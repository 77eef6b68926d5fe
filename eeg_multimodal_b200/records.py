"""Epoch-level evaluation records in the reference's format.

Reference: past_acc.py:218-250 -- per-batch loss/accuracy averaged UNWEIGHTED over batches
(`epoch_acc_val / sample_size_val`, the last batch of the 601-row split has one sample), binary
F1 over all predictions (`f1_score(prediction_all, label_all)`: arguments swapped, which leaves the
binary F1 unchanged), record appended to whole_record.txt, best kept in best_record.txt and the
state_dict saved when F1 improves, best initialised to 0.5 (past_acc.py:182,242-250).
Format sample: model_dict/newfrac_1.0eps/best_record.txt:1-6.
"""
from __future__ import annotations

import os

import torch


def binary_f1(pred: torch.Tensor, label: torch.Tensor) -> float:
    """sklearn.metrics.f1_score(pred, label) for {0,1} labels, positive class 1, computed from the
    int64 tensors the CE kernel already produced (zero_division -> 0.0 like sklearn's default)."""
    pred, label = pred.reshape(-1).long(), label.reshape(-1).long()
    tp = int(((pred == 1) & (label == 1)).sum())
    fp = int(((pred == 1) & (label == 0)).sum())
    fn = int(((pred == 0) & (label == 1)).sum())
    denom = 2 * tp + fp + fn
    return 0.0 if denom == 0 else 2.0 * tp / denom


def format_record(epoch, train_loss, train_acc, val_loss, val_acc, f1) -> str:
    """The f-string of past_acc.py:232-237, byte for byte (8-space indent, `: .3f` formats)."""
    return f'''Epochs: {epoch}
        | Train Loss: {train_loss: .3f}
        | Train Accuracy: {train_acc: .3f}
        | Val Loss: {val_loss: .3f}
        | Val Accuracy: {val_acc: .3f}
        | f_1 Score: {f1: .3f}\n'''


class EpochMeter:
    """Accumulates per-batch loss / accuracy the way the reference does (sum of batch means divided
    by the number of batches) plus all predictions and labels for the F1."""

    def __init__(self):
        self.loss_sum = self.acc_sum = 0.0
        self.batches = 0
        self.preds, self.labels = [], []

    def update(self, loss: float, acc: float, pred=None, label=None):
        self.loss_sum += float(loss)
        self.acc_sum += float(acc)
        self.batches += 1
        if pred is not None:
            self.preds.append(pred.reshape(-1).cpu())
            self.labels.append(label.reshape(-1).cpu())

    @property
    def loss(self):
        return self.loss_sum / max(1, self.batches)

    @property
    def acc(self):
        return self.acc_sum / max(1, self.batches)

    def f1(self):
        return binary_f1(torch.cat(self.preds), torch.cat(self.labels)) if self.preds else 0.0


class RecordWriter:
    """whole_record.txt / best_record.txt / best_f1.pickle under `root/<suffix>` (past_acc.py:164-168)."""

    def __init__(self, root: str, suffix: str = ""):
        self.dir = os.path.join(root, suffix)
        os.makedirs(self.dir, exist_ok=True)
        self.whole = os.path.join(self.dir, "whole_record.txt")
        self.best = os.path.join(self.dir, "best_record.txt")
        self.ckpt = os.path.join(self.dir, "best_f1.pickle")  # a torch zip despite the name, like the reference
        self.f1_best = 0.5

    def epoch_end(self, epoch, train: EpochMeter, val: EpochMeter, state_dict_fn=None) -> bool:
        f1 = val.f1()
        rec = format_record(epoch, train.loss, train.acc, val.loss, val.acc, f1)
        with open(self.whole, "a") as f:
            f.write(rec)
        improved = f1 > self.f1_best
        if improved:
            if state_dict_fn is not None:
                torch.save(state_dict_fn(), self.ckpt)
            self.f1_best = f1
            with open(self.best, "w") as f:
                f.write(rec)
        return improved
