"""Host-side score table of `eval` (utils.py:111-156 `scoring`): the same keys and definitions, computed with
scikit-learn once per pass on the collected predictions.  Not on the per-batch path."""
from __future__ import annotations

import torch


def scoring(y_true: torch.Tensor, y_pred: torch.Tensor, y_pred_confidence: torch.Tensor, is_regression: bool = False,
            is_multiclass: bool = False, require_sigmoid: bool = True, labels=None) -> dict:
    if is_regression:
        return {"mae": torch.nn.functional.l1_loss(y_pred, y_true).item(),
                "mse": torch.nn.functional.mse_loss(y_pred, y_true).item()}
    from sklearn import metrics as M
    s = {}
    if is_multiclass:
        from sklearn.preprocessing import OneHotEncoder
        probs = y_pred_confidence.softmax(-1)
        onehot = OneHotEncoder().fit_transform(y_true.reshape(-1, 1)).toarray()
        hard = probs.argmax(-1)
        s["auc"] = M.roc_auc_score(onehot, probs, multi_class="ovr", average="macro")
        s["f1_score"] = M.f1_score(y_true, hard, average="macro")
        s["accuracy"] = M.accuracy_score(y_true, hard)
        s["ap"] = M.average_precision_score(onehot, probs, average="macro")
        s["loss"] = None                                     # `eval` overwrites it with criterion(confidence, y_true)
    else:
        conf = y_pred_confidence.sigmoid() if require_sigmoid else y_pred_confidence
        s["auc"] = M.roc_auc_score(y_true, y_pred_confidence.sigmoid())
        s["f1_score"] = M.f1_score(y_true, y_pred)
        s["accuracy"] = M.accuracy_score(y_true, y_pred)
        s["ap"] = M.average_precision_score(y_true, conf)
        s["loss"] = torch.nn.BCEWithLogitsLoss()(y_pred.float(), y_true.float()).item()
    s["confusion_matrix"] = M.confusion_matrix(y_true, y_pred, labels=labels)
    return s
