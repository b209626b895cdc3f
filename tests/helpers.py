"""Shared helpers for the parity tests (fixture parsing, tolerances)."""
import os
import re

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def loss_cases():
    """-> list of (loss_key, case, dict(scores, labels, loss, grad[, perm]))."""
    z = load("losses.npz")
    groups = {}
    for k in z.files:
        head, case, field = k.split("__")
        groups.setdefault((head, case), {})[field] = z[k]
    return [(h, c, v) for (h, c), v in sorted(groups.items())]


def parse_loss_key(head):
    """'LambdaLoss_NDCG_Loss2++_k5_unsorted' -> (name, params, presort)."""
    presort = True
    if head.endswith("_unsorted"):
        presort, head = False, head[: -len("_unsorted")]
    if head.endswith("_saturated"):
        head = head[: -len("_saturated")]
    name = head.split("_")[0]
    params = {}
    m = re.search(r"sigma([0-9.]+)", head)
    if m:
        params["sigma"] = float(m.group(1))
    m = re.search(r"alpha([0-9.]+)", head)
    if m:
        params["alpha"] = float(m.group(1))
    if name == "LambdaLoss":
        m = re.match(r"LambdaLoss_(NDCG_Loss(?:1|2\+\+|2))_k(\d+)", head)
        params.update(loss_type=m.group(1), k=int(m.group(2)), sigma=1.0, mu=5.0)
    return name, params, presort


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / denom)


def sibling_cases():
    """tests/golden/siblings.npz -> list of (loss_key, case, dict) for RankMSE / RankCosine / STListNet / SoftRank."""
    z = load("siblings.npz")
    groups = {}
    for k in z.files:
        head, case, field = k.split("__")
        if head in ("sinkstep", "sinkhorn"):
            continue
        groups.setdefault((head, case), {})[field] = z[k]
    return [(h, c, v) for (h, c), v in sorted(groups.items())]


def sinkhorn_cases(kind):
    z = load("siblings.npz")
    groups = {}
    for k in z.files:
        head, case, field = k.split("__")
        if head == kind:
            groups.setdefault(case, {})[field] = z[k]
    return sorted(groups.items())


def parse_sibling_key(head):
    """'SoftRank_delta2.0_kNone' -> ('SoftRank', dict(delta=2.0, top_k=None))."""
    name = head.split("_")[0]
    params = {}
    m = re.search(r"_T([0-9.]+)", head)
    if m:
        params["temperature"] = float(m.group(1))
    m = re.search(r"delta([0-9.]+)_k(\w+)", head)
    if m:
        params["delta"] = float(m.group(1))
        params["top_k"] = None if m.group(2) == "None" else int(m.group(2))
    return name, params


def sampled(arr, limit=16384, target=8192):
    """Flattened tensor, or every k-th element of it when it has more than ``limit`` elements (k = size // target) --
    the storage rule of tests/golden/make_golden_r2.py for large gradients / weights."""
    flat = np.asarray(arr).reshape(-1)
    if flat.size <= limit:
        return flat.copy()
    return flat[:: flat.size // target].copy()
