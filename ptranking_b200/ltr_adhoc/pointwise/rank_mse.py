"""RankMSE (mirror of ptranking/ltr_adhoc/pointwise/rank_mse.py:24-43)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ... import ops


class RankMSE(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, gpu=False, device=None):
        super().__init__(id='RankMSE', sf_para_dict=sf_para_dict, gpu=gpu, device=device)

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """Mean over the batch of the per-query summed squared error (rank_mse.py:20-21)."""
        batch_loss = ops.rank_loss('RankMSE', batch_preds, batch_std_labels, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
