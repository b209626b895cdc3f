"""LambdaRank (mirror of ptranking/ltr_adhoc/listwise/lambdarank.py:18-62)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ...base.ranker import _is_multilabel
from ... import ops


class LambdaRank(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, model_para_dict=None, gpu=False, device=None):
        super().__init__(id='LambdaRank', sf_para_dict=sf_para_dict, gpu=gpu, device=device)
        self.sigma = model_para_dict['sigma']

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """Delta-nDCG weighted pairwise BCE on the predicted order; same preconditions as the
        reference (lambdarank.py:34-36): MultiLabel labels, presorted descending."""
        assert 'label_type' in kwargs and _is_multilabel(kwargs['label_type'])
        assert 'presort' in kwargs and kwargs['presort'] is True
        batch_loss = ops.rank_loss('LambdaRank', batch_preds, batch_std_labels, sigma=self.sigma, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
