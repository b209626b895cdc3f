"""A few training steps at a bench workload, for ncu (never a bench value).
    python tools/profile_step.py [steps] [B] [config a|b|c|d|e]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from ptranking_b200 import LABEL_TYPE

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 0
config = sys.argv[3] if len(sys.argv) > 3 else "b"
args = argparse.Namespace(config=config, batch=B, cpu_batch=0, docs=256, enc_layers=int(os.environ.get("ENC_LAYERS", "6")))
cfg = bench.make_config(args)
torch.manual_seed(137)
rng = np.random.default_rng(137)
r = bench.build_ranker(cfg, "cuda:0")
X, y = bench.synth_batch(rng, cfg["B"], cfg["n"], cfg["F"], cfg["labels"])
X, y = X.cuda(), y.cuda()
for i in range(steps):
    r.train_op(X, y, presort=True, label_type=LABEL_TYPE.MultiLabel, epoch_k=1)
torch.cuda.synchronize()
print("done", steps, cfg["key"])
