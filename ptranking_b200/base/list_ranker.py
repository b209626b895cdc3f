"""List (multi-head self-attention) ranker -- mirror of ptranking/base/list_ranker.py:280-402.
The fused encoder kernels are filled in by csrc/listsf.cu (see DESIGN.md for status)."""
from __future__ import annotations

from .ranker import NeuralRanker


class ListNeuralRanker(NeuralRanker):
    def __init__(self, id='ListNeuralRanker', sf_para_dict=None, weight_decay=1e-3, gpu=False, device=None):
        super().__init__(id=id, sf_para_dict=sf_para_dict, weight_decay=weight_decay, gpu=gpu, device=device)
        self.encoder_type = self.sf_para_dict[self.sf_para_dict['sf_id']]['encoder_type']

    def init(self):
        raise NotImplementedError("listsf scorer kernels are not built yet")
