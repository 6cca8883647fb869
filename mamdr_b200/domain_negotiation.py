"""Domain Negotiation -- mirrors ``/root/reference/model_zoo/domain_negotiation.py`` (Alg. 1):
shuffle the domains, model <- theta, one full pass per domain with the single Adam, then
theta <- theta + beta * (model - theta).
"""
from .engine import _ptr
from .maml import MAML


class DomainNegotiation(MAML):
    def __init__(self, base_model):
        super(DomainNegotiation, self).__init__(base_model)

    def prepare(self):
        self._get_model_meta_parms()                          # :27
        self.meta_weights = self._get_meta_weights()          # :29
        self.model.reset_optimizer()                          # :31 global_variables_initializer
        self.meta_sequence = self.build_meta_data_split()     # :33

    def train(self):
        self.log("Start Domain Negotiation on model: {}".format(self.model_config['name']))
        self.prepare()
        for epoch in range(self.train_config['epoch']):
            self.log("Epoch: {}".format(epoch), "-" * 30)
            self.train_epoch(epoch)
            if epoch % self.train_config['val_every_step'] == 0:   # :95-109
                val_avg_loss, val_avg_auc, val_domain_loss, val_domain_auc = self.val()
                val_metric = val_domain_auc[self.train_config['target_domain']] \
                    if self.train_config['target_domain'] >= 0 else val_avg_auc
                if self.early_stop_step(val_metric):
                    break
                self.log("Test Result: ")
                self.val_and_test("test")

    def train_epoch(self, epoch=0):
        """One DN meta-step (:41-93)."""
        tc = self.train_config
        target = tc['target_domain']
        if tc['shuffle_sequence']:                            # :41-42
            self.meta_sequence = self.schedule.shuffle_sequence(self.meta_sequence)
        train_sequence = list(self.meta_sequence) + ([target] if target >= 0 else [])   # :44-47
        self.stage_epoch_orders(train_sequence + ([target] if target >= 0 else []))
        # the whole meta-step is recorded and runs as ONE persistent launch in the tcgen05 modes (engine.program)
        with self.model.program(self.b200_config.get('program', True)):
            self._set_model_meta_parms(self.meta_weights)         # :50
            for idx in train_sequence:                            # :53-84
                d = self.dataset.train_dataset[idx]
                for m in self.model.stateful_metric_functions:    # :56-57
                    m.reset_states()
                train_step = d['n_step']
                if tc['meta_train_step'] > 0 and idx != target:   # :67
                    train_step = min(train_step, tc['meta_train_step'])
                # per-pass loss / AUC prints of :80-84 would force a host sync per pass; the per-batch
                # losses stay on the device (fit_pass returns them) and are read only on request
                self.last_pass_losses = self.run_train_pass(idx, train_step)
            self._update_meta_weight(self.meta_weights)           # :87
            # :88 _set_model_meta_parms(meta_weights) is fused into the update (model_out)
            if target >= 0:                                       # :89-93 model.fit(target_iter, steps_per_epoch=target_step)
                for m in self.model.stateful_metric_functions:
                    m.reset_states()
                self.last_pass_losses = self.run_train_pass(target)

    def _update_meta_weight(self, old_vars):
        """:118-123  old += (new - old) * meta_learning_rate, and the model is reloaded with it (:88)."""
        m = self.model
        beta = self.train_config['meta_learning_rate']
        for n, (theta, model) in self._ranges(old_vars.flat, m.params):
            m.ctx.call("mamdr_dn_update", _ptr(theta), _ptr(model), beta, n, _ptr(model), m.stream)
            m.ctx.launches += 1

    def build_meta_data_split(self):
        """:125-146 -- the meta sequence (data iterators are the device-resident column stores)."""
        target = self.train_config['target_domain']           # :138-140 the target domain is not part of the meta sequence
        meta_sequence = [k for k in self.dataset.train_dataset.keys() if not (target >= 0 and k == target)]
        ms = self.train_config.get('meta_sequence')
        if isinstance(ms, list):
            if len(ms) != len(meta_sequence):
                raise ValueError("All the domains must be given in the sequence")
            meta_sequence = list(ms)
        return meta_sequence
