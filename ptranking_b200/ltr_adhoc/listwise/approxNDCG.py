"""ApproxNDCG (mirror of ptranking/ltr_adhoc/listwise/approxNDCG.py:67-107)."""
from ...base.adhoc_ranker import AdhocNeuralRanker
from ...base.ranker import _is_multilabel
from ... import ops


class ApproxNDCG(AdhocNeuralRanker):
    def __init__(self, sf_para_dict=None, model_para_dict=None, gpu=False, device=None):
        super().__init__(id='ApproxNDCG', sf_para_dict=sf_para_dict, gpu=gpu, device=device)
        self.alpha = model_para_dict['alpha']
        # True reproduces the reference's [B]/[B,1] broadcast (approxNDCG.py:58-61); False is the
        # per-query normalisation the paper describes.
        self.batch_coupled = bool(model_para_dict.get('batch_coupled', True))

    def uniform_eval_setting(self, **kwargs):
        eval_dict = kwargs['eval_dict']
        if eval_dict["do_validation"] and not eval_dict['vali_metric'] == 'nDCG':
            eval_dict['vali_metric'] = "nDCG"

    def custom_loss_function(self, batch_preds, batch_std_labels, **kwargs):
        assert _is_multilabel(kwargs['label_type'])
        presort = bool(kwargs.get('presort', False))
        batch_loss = ops.rank_loss('ApproxNDCG', batch_preds, batch_std_labels, alpha=self.alpha,
                                   presort=presort, batch_coupled=self.batch_coupled, **self.ragged_kwargs(kwargs))
        self.backward_and_step(batch_loss)
        return batch_loss
