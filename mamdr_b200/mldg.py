"""MLDG -- mirrors ``/root/reference/model_zoo/mldg.py``: the MAML loop with the inner Adam
steps replaced by gradient accumulation.  Per domain: model <- theta; the gradients of the meta-train batches are ACCUMULATED
at theta (:94-96); ONE meta-Adam apply moves the live model along them without clearing the accumulators (:107-108); the
gradients of the meta-val batches at the moved weights are added to the same accumulators (:112-114); then model <- theta and
the meta Adam applies the sum and clears it (:119-120) -- or, for ``batch`` names, once per epoch (:123-125).  Everything
else (meta-parameter selection, the accumulating function, the data split) is MAML's (``mldg.py:157-366`` == ``maml.py:153-363``).
"""
from .maml import MAML


class MLDG(MAML):
    def _inner_loop(self, idx, d):
        tc = self.train_config
        train_step, meta_val_step = d['train_step'], d['meta_val_step']
        if tc['meta_train_step'] > 0:                      # :88-90
            train_step = min(train_step, tc['meta_train_step'])
            meta_val_step = min(meta_val_step, tc['meta_train_step'])
        self.meta_train_pass(d['train_iter'], train_step)  # :92-96  accumulate at theta
        self.meta_parms_update_step()                      # :107-108 K.get_session().run(self.meta_parms_update_step)
        self.meta_train_pass(d['meta_iter'], meta_val_step)   # :112-114 accumulate at the moved weights
