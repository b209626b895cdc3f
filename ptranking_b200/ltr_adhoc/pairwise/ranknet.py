"""RankNet (mirror of ptranking/ltr_adhoc/pairwise/ranknet.py:18-42)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ... import ops


class RankNet(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, model_para_dict=None, gpu=False, device=None):
        super().__init__(id='RankNet', sf_para_dict=sf_para_dict, gpu=gpu, device=device)
        self.sigma = model_para_dict['sigma']

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """BCE over all pairs i<j of sigmoid(sigma (s_i - s_j)) vs 1/2 (1 + sign(y_i - y_j)); one fused kernel."""
        batch_loss = ops.rank_loss('RankNet', batch_preds, batch_std_labels, sigma=self.sigma, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
