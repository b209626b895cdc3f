"""RankCosine (mirror of ptranking/ltr_adhoc/listwise/rank_cosine.py:17-40)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ... import ops


class RankCosine(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, gpu=False, device=None):
        super().__init__(id='RankCosine', sf_para_dict=sf_para_dict, gpu=gpu, device=device)

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """sum_q (1 - cos(preds_q, labels_q)) / 0.5 (rank_cosine.py:33)."""
        batch_loss = ops.rank_loss('RankCosine', batch_preds, batch_std_labels, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
