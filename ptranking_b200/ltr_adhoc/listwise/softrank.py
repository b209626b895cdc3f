"""SoftRank (mirror of ptranking/ltr_adhoc/listwise/softrank.py:21-78)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ...base.ranker import _is_multilabel
from ... import ops


class SoftRank(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, model_para_dict=None, gpu=False, device=None):
        super().__init__(id='SoftRank', sf_para_dict=sf_para_dict, gpu=gpu, device=device)
        self.delta = float(model_para_dict['delta'])
        self.top_k = model_para_dict['top_k']
        self.metric = model_para_dict['metric']

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """Negative expected nDCG under Gaussian score noise; same preconditions as the reference (softrank.py:40-42)."""
        assert 'presort' in kwargs and kwargs['presort'] is True
        assert 'nDCG' == self.metric
        assert _is_multilabel(kwargs['label_type'])
        batch_loss = ops.rank_loss('SoftRank', batch_preds, batch_std_labels, delta=self.delta, top_k=self.top_k, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
