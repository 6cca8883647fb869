"""PCGrad -- mirrors ``/root/reference/model_zoo/pcgrad.py``.  Per epoch and per domain (in a
shuffled order): the gradients of ALL its mini-batches are accumulated at the live weights (:87-96); ``sample_num`` support
domains are drawn (:109-111); each support domain's accumulated gradient is projected against the running sum and added to it
(``PCGrad`` :152-160 -- host numpy in the reference, ``mamdr_pcgrad_project`` on the device here); the meta Adam applies the
sum to the live model (:126-127).  The model is never reset to a theta: the meta optimizer is the only thing that moves it.
"""
from .engine import _ptr
from .maml import MAML


class PCGrad(MAML):
    def build_meta_data_split(self):
        """:324-330 -- every domain's own training dataset (its shuffle and batching), no meta split."""
        return {idx: {"train_iter": d['data'], "train_step": d['n_step']} for idx, d in self.dataset.train_dataset.items()}

    def _init_iter(self, data):
        data.set_order(self.schedule.batch_order(data.domain, data.n_data))

    def prepare(self):
        super(PCGrad, self).prepare()
        import torch
        self._final = torch.zeros_like(self.accum_grads)

    def project(self, final, aux):
        """``self.PCGrad(final_grads, current_grads, aux_grads)`` (:152-160); ``final_grads`` IS ``current_grads`` (:104)."""
        m = self.model
        for p in self.model_meta_parms:
            cols = int(p.shape[-1])
            rows = p.numel // cols
            m.ctx.call("mamdr_pcgrad_project", _ptr(final[p.offset:p.offset + p.numel]), _ptr(aux[p.offset:p.offset + p.numel]),
                       rows, cols, m.stream)
            m.ctx.launches += 1

    def finish_epoch(self):
        pass

    def domain_step(self, idx):
        tc = self.train_config
        m = self.model
        d = self.meta_data_split[idx]
        for metric in m.stateful_metric_functions:                               # :71-72
            metric.reset_states()
        self._init_iter(d['train_iter'])                                         # :79
        train_step = d['train_step']
        if tc['meta_train_step'] > 0:                                            # :82-83
            train_step = min(train_step, tc['meta_train_step'])
        self.clear_grads()                                                       # :86
        self.meta_train_pass(d['train_iter'], train_step)                        # :88-91
        for n, (fin, acc) in self._ranges(self._final, self.accum_grads):        # :103-104 current_grads = final_grads
            m.ctx.call("mamdr_copy", _ptr(fin), _ptr(acc), n, m.stream)
            m.ctx.launches += 1
        candidates = list(self.train_sequence)                                   # :107-109
        candidates.remove(idx)
        for aux_idx in self.schedule.sample_support(candidates, tc['sample_num']):
            aux_d = self.meta_data_split[aux_idx]
            self._init_iter(aux_d['train_iter'])                                 # :115
            self.clear_grads()                                                   # :118
            self.meta_train_pass(aux_d['train_iter'], aux_d['train_step'])       # :120-121 (no meta_train_step cap here)
            self.project(self._final, self.accum_grads)                          # :123-124
        for n, (acc, fin) in self._ranges(self.accum_grads, self._final):        # :127 set_accum_grads(final_grads)
            m.ctx.call("mamdr_copy", _ptr(acc), _ptr(fin), n, m.stream)
            m.ctx.launches += 1
        self._meta_train_step()                                                  # :128
