"""Golden vectors for P / AP / nERR from the UNMODIFIED reference (run in the authoring container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_metrics2.py
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, os.environ.get("PTRANKING_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
from ptranking.metric.adhoc.adhoc_metric import (torch_ap_at_ks, torch_ndcg_at_ks, torch_nerr_at_ks,  # noqa: E402
                                                 torch_precision_at_ks)

out = {}
# the reference's own known answers (testing/metric/testing_metric.py:20-60)
kats = [
    ("ap1", [1., 0., 1., 0., 1.], [1., 1., 1., 1., 1.], [1, 3, 5], [1.0000, 0.5556, 0.4533]),
    ("ap2", [1., 0., 1., 0., 1.], [1., 1., 1., 0., 0.], [1, 3, 5], [1.0000, 0.5556, 0.7556]),
    ("ap3", [1., 1., 0., 1., 0., 0., 1.], [1., 1., 1., 1., 0., 0., 0.], [1, 2, 3, 5, 7], [1.0000, 1.0000, 0.6667, 0.6875, 0.8304]),
]
for name, sys_l, std_l, ks, expect in kats:
    s, t = torch.tensor([sys_l]), torch.tensor([std_l])
    out[f"{name}__sys"] = s.numpy(); out[f"{name}__std"] = t.numpy(); out[f"{name}__ks"] = np.array(ks)
    out[f"{name}__ap"] = torch_ap_at_ks(s, t, ks=ks).numpy(); out[f"{name}__expect4dp"] = np.array(expect)
s, t = torch.tensor([[3., 2., 4.]]), torch.tensor([[4., 3., 2.]])
out["nerr__sys"] = s.numpy(); out["nerr__std"] = t.numpy(); out["nerr__ks"] = np.array([1, 2, 3])
out["nerr__val"] = torch_nerr_at_ks(s, t, ks=[1, 2, 3]).numpy(); out["nerr__expect4dp"] = np.array([0.4667, 0.5154, 0.6640])

rng = np.random.default_rng(137)
P5 = np.array([1940952, 1225770, 504958, 69010, 30435], dtype=np.float64); P5 /= P5.sum()
for (B, n) in [(5, 50), (3, 256), (2, 7), (2, 1024)]:
    y = rng.choice(5, size=(B, n), p=P5).astype(np.float32)
    y[:, 0] = np.maximum(y[:, 0], 1)
    y = -np.sort(-y, axis=1)
    s = rng.standard_normal((B, n)).astype(np.float32)
    ks = [1, 3, 5, 10, 20, 50]
    ts, ty = torch.from_numpy(s), torch.from_numpy(y)
    _, idx = torch.sort(ts, dim=1, descending=True)
    sys_r = torch.gather(ty, 1, idx)
    key = f"B{B}_n{n}"
    out[key + "__scores"] = s; out[key + "__labels"] = y; out[key + "__ks"] = np.array(ks)
    out[key + "__ndcg"] = torch_ndcg_at_ks(sys_r, ty, ks=ks).numpy()
    out[key + "__nerr4"] = torch_nerr_at_ks(sys_r, ty, ks=ks, max_label=4.0).numpy()
    out[key + "__nerrNone"] = torch_nerr_at_ks(sys_r, ty, ks=ks, max_label=None).numpy()
    out[key + "__ap"] = torch_ap_at_ks(sys_r, ty, ks=ks).numpy()
    out[key + "__p"] = torch_precision_at_ks(sys_r, ks=ks).numpy()
np.savez_compressed(os.path.join(HERE, "metrics2.npz"), **out)
print("metrics2.npz:", len(out), "arrays")
