"""Reptile -- mirrors ``/root/reference/model_zoo/reptile.py``: a sibling of Domain
Negotiation on the same kernels.  Per epoch: shuffle the domains; for every domain: model <- theta, one pass with the
single Adam, then theta <- theta + beta (model - theta) (:127-132) -- or, for ``batch`` names, accumulate the deltas
(:134-137) and apply their sum once per epoch (:139-142).  Unlike DN the model is RESET to theta before every domain.
"""
import torch

from .engine import _ptr
from .maml import MAML


class Reptile(MAML):
    def __init__(self, base_model):
        super(Reptile, self).__init__(base_model)

    def prepare(self):
        """reptile.py:26-35 -- meta parameter set, theta, zeroed accumulators, fresh optimizer slots."""
        self._get_model_meta_parms()                           # :27
        self.meta_weights = self._get_meta_weights()           # :29
        self.accum_grads = torch.zeros_like(self.meta_weights.flat)   # :31
        self._zeros = torch.zeros_like(self.meta_weights.flat)
        self.model.reset_optimizer()                           # :33 global_variables_initializer
        self.train_sequence = list(range(self.n_domain))       # :36

    def train(self):
        self.log("Start reptile on model: {}".format(self.model_config['name']))
        self.prepare()
        for epoch in range(self.train_config['epoch']):
            self.log("Epoch: {}".format(epoch), "-" * 30)
            self.train_epoch(epoch)
            if epoch % self.train_config['val_every_step'] == 0:   # :105-118
                val_avg_loss, val_avg_auc, val_domain_loss, val_domain_auc = self.val()
                val_metric = val_domain_auc[self.train_config['target_domain']] \
                    if self.train_config['target_domain'] >= 0 else val_avg_auc     # :112-113
                if self.early_stop_step(val_metric):
                    break
                self.log("Test Result: ")
                self.val_and_test("test")

    def train_epoch(self, epoch=0):
        """One Reptile epoch (:45-97)."""
        tc = self.train_config
        target = tc['target_domain']
        batch = "batch" in self.model_config['name']
        m = self.model
        beta = tc['meta_learning_rate']
        self.train_sequence = self.schedule.shuffle_sequence(self.train_sequence)     # :46
        inner = [idx for idx in self.train_sequence if not (target >= 0 and idx == target)]          # :47-48
        passes = []
        for idx in inner:
            passes += [idx] + ([target] if target >= 0 else [])
        self.stage_epoch_orders(passes + ([target] if target >= 0 else []))
        with m.program(self.b200_config.get('program', True)):
            self._set_model_meta_parms(self.meta_weights)        # :57 (first domain; later ones are fused into the update)
            for idx in inner:
                d = self.dataset.train_dataset[idx]
                for metric in m.stateful_metric_functions:       # :53-54
                    metric.reset_states()
                train_step = d['n_step']
                if tc['meta_train_step'] > 0:
                    train_step = min(train_step, tc['meta_train_step'])
                self.last_pass_losses = self.run_train_pass(idx, train_step)
                if target >= 0:                                  # :82-85 one step on the target domain before the update
                    self.run_train_pass(target, 1)
                if batch:                                        # :90-91  accum += model - theta ; then model <- theta (:57)
                    for n, (acc, model, theta) in self._ranges(self.accum_grads, m.params, self.meta_weights.flat):
                        m.ctx.call("mamdr_axpy_diff", _ptr(acc), _ptr(model), _ptr(theta), 1.0, n, m.stream)
                        m.ctx.call("mamdr_copy", _ptr(model), _ptr(theta), n, m.stream)
                        m.ctx.launches += 2
                else:                                            # :92-93  theta += (model - theta) * beta ; model <- theta (:57 / :99)
                    for n, (theta, model) in self._ranges(self.meta_weights.flat, m.params):
                        m.ctx.call("mamdr_dn_update", _ptr(theta), _ptr(model), beta, n, _ptr(model), m.stream)
                        m.ctx.launches += 1
            if batch:                                            # :96-97, :139-142  theta += accum * beta ; accum <- 0 ; model <- theta (:99)
                for n, (theta, acc, zeros, model) in self._ranges(self.meta_weights.flat, self.accum_grads, self._zeros, m.params):
                    m.ctx.call("mamdr_axpy_diff", _ptr(theta), _ptr(acc), _ptr(zeros), beta, n, m.stream)
                    m.ctx.call("mamdr_copy", _ptr(acc), _ptr(zeros), n, m.stream)
                    m.ctx.call("mamdr_copy", _ptr(model), _ptr(theta), n, m.stream)
                    m.ctx.launches += 3
            if target >= 0:                                      # :98-102 model.fit(target_iter, steps_per_epoch=target_step)
                for metric in m.stateful_metric_functions:
                    metric.reset_states()
                self.last_pass_losses = self.run_train_pass(target)
