"""LambdaLoss (mirror of ptranking/ltr_adhoc/listwise/lambdaloss.py:61-138)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ...base.ranker import _is_multilabel
from ... import ops

LAMBDALOSS_TYPE = ['NDCG_Loss1', 'NDCG_Loss2', 'NDCG_Loss2++']


class LambdaLoss(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, model_para_dict=None, gpu=False, device=None):
        super().__init__(id='LambdaLoss', sf_para_dict=sf_para_dict, gpu=gpu, device=device)
        self.lambdaloss_dict = model_para_dict
        self.k, self.sigma, self.loss_type = model_para_dict['k'], model_para_dict['sigma'], model_para_dict['loss_type']
        assert self.loss_type in LAMBDALOSS_TYPE
        self.mu = model_para_dict['mu'] if 'NDCG_Loss2++' == self.loss_type else 0.0

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        assert _is_multilabel(kwargs['label_type'])
        presort = bool(kwargs.get('presort', False))
        batch_loss = ops.rank_loss('LambdaLoss', batch_preds, batch_std_labels, k=self.k, sigma=self.sigma,
                                   mu=self.mu, loss_type=self.loss_type, presort=presort, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
