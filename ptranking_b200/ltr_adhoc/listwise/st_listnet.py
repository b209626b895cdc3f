"""STListNet (mirror of ptranking/ltr_adhoc/listwise/st_listnet.py:22-57)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ... import ops


class STListNet(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, model_para_dict=None, gpu=False, device=None):
        super().__init__(id='STListNet', sf_para_dict=sf_para_dict, gpu=gpu, device=device)
        self.temperature = model_para_dict['temperature']

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        """ListNet's top-1 cross entropy on Gumbel-perturbed scores (st_listnet.py:41-49).  The uniform draw comes from
        the kernel's counter-based generator (the reference calls torch.rand); ``unif=`` injects one (parity tests)."""
        batch_loss = ops.rank_loss('STListNet', batch_preds, batch_std_labels, temperature=self.temperature,
                                   unif=kwargs.get('unif'), **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
