"""AdhocNeuralRanker: the class every ltr_adhoc loss subclasses (mirror of
ptranking/base/adhoc_ranker.py:7-87) -- picks the pointwise or the list scorer by ``sf_id``."""
from __future__ import annotations

from .list_ranker import ListNeuralRanker
from .point_ranker import PointNeuralRanker


class AdhocNeuralRanker(PointNeuralRanker, ListNeuralRanker):
    def __init__(self, id='AdhocNeuralRanker', sf_para_dict=None, weight_decay=1e-3, gpu=False, device=None):
        self.id = id
        self.gpu, self.device = gpu, device
        self.sf_para_dict = sf_para_dict
        self.sf_id = sf_para_dict['sf_id']
        assert self.sf_id in ['pointsf', 'listsf']
        self.opt, self.lr = sf_para_dict['opt'], sf_para_dict['lr']
        self.weight_decay = weight_decay
        self.stop_check_freq = 10
        if 'listsf' == self.sf_id:
            self.encoder_type = self.sf_para_dict[self.sf_para_dict['sf_id']]['encoder_type']
        self._require_cuda()

    def _base(self):
        return PointNeuralRanker if 'pointsf' == self.sf_id else ListNeuralRanker

    def init(self):
        self._base().init(self)

    def get_parameters(self):
        return self._base().get_parameters(self)

    def forward(self, batch_q_doc_vectors):
        return self._base().forward(self, batch_q_doc_vectors)

    def forward_ragged(self, flat_q_doc_vectors, offsets, max_len, buckets=None):
        return self._base().forward_ragged(self, flat_q_doc_vectors, offsets, max_len, buckets=buckets)

    def eval_mode(self):
        self._base().eval_mode(self)

    def train_mode(self):
        self._base().train_mode(self)

    def save(self, dir, name):
        self._base().save(self, dir=dir, name=name)

    def load(self, file_model, device=None, **kwargs):
        # the reference forwards device= to ListNeuralRanker.load, which does not accept it
        # (adhoc_ranker.py:81 / list_ranker.py:398); both scorers accept it here.
        self._base().load(self, file_model=file_model, device=device)

    def get_tl_af(self):
        key = 'TL_AF' if 'pointsf' == self.sf_id else 'AF'
        return self.sf_para_dict[self.sf_para_dict['sf_id']][key]
