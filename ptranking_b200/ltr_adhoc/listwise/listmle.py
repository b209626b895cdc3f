"""ListMLE (mirror of ptranking/ltr_adhoc/listwise/listmle.py:65-104)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ... import ops


class ListMLE(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, model_para_dict=None, gpu=False, device=None):
        super().__init__(id='ListMLE', sf_para_dict=sf_para_dict, gpu=gpu, device=device)

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """Plackett-Luce likelihood of the ideal ordering; ties in the labels are re-shuffled on every
        call like the reference (listmle.py:81) -- by a Philox kernel instead of B host-side randperms.
        ``perm=`` (int [B,n]) injects a fixed ordering (parity tests)."""
        perm = kwargs.get('perm')
        if perm is None:
            perm = ops.shuffle_ties_perm(batch_std_labels, **self.ragged_kwargs(kwargs))
        batch_loss = ops.rank_loss('ListMLE', batch_preds, batch_std_labels, perm=perm, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
