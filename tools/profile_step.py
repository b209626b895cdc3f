"""A few LambdaRank train steps at the bench workload, for ncu (never a bench value)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import ptranking_b200
from ptranking_b200 import LABEL_TYPE

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(137)
rng = np.random.default_rng(137)
r = ptranking_b200.LambdaRank(sf_para_dict=bench.default_sf(), model_para_dict=dict(model_id="LambdaRank", sigma=1.0),
                              gpu=True, device="cuda:0")
r.init()
r.train_mode()
X, y = bench.synth_batch(rng, B)
X, y = X.cuda(), y.cuda()
for i in range(steps):
    r.train_op(X, y, presort=True, label_type=LABEL_TYPE.MultiLabel, epoch_k=1)
torch.cuda.synchronize()
print("done", steps)
