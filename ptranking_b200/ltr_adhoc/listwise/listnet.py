"""ListNet (mirror of ptranking/ltr_adhoc/listwise/listnet.py:14-45)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ... import ops


class ListNet(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, gpu=False, device=None):
        super().__init__(id='ListNet', sf_para_dict=sf_para_dict, gpu=gpu, device=device)

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """Top-1 ListNet: cross entropy of softmax(scores) against softmax(labels)."""
        batch_loss = ops.rank_loss('ListNet', batch_preds, batch_std_labels, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
