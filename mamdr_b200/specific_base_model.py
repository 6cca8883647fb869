"""Shared + domain-specific parameter bookkeeping -- mirrors
``/root/reference/model_zoo/specific_base_model.py``: best-snapshot early stopping (:44-62),
``val_and_test`` with merged weights per domain (:64-97), the per-domain SGD finetune stage
(:99-162), ``_merge_weights`` (:164-172) and ``init_layer`` (:174-178).
"""
import os
import os.path as osp

import torch

from . import _lib
from .base_model import keras_fit_with_callbacks
from .engine import WEIGHT_EXT, _ptr
from .maml import MAML, MetaWeights


class SpecificBase(MAML):
    def __init__(self, base_model):
        super(SpecificBase, self).__init__(base_model)

    def build_meta_data_split(self):
        """:20-42"""
        target = self.train_config['target_domain']           # :33-36 the target domain is skipped (it is only evaluated)
        meta_sequence = [k for k in self.dataset.train_dataset.keys() if not (target >= 0 and k == target)]
        ms = self.train_config.get('meta_sequence')
        if isinstance(ms, list):
            if len(ms) != len(meta_sequence):
                raise ValueError("All the domains must be given in the sequence")
            meta_sequence = list(ms)
        return meta_sequence

    def _merge_method(self):
        mm = self.train_config['merged_method']
        if mm == 'plus':
            return _lib.MERGE_PLUS
        if mm == 'times':
            return _lib.MERGE_TIMES
        raise ValueError("merged_method must be 'plus' or 'times', not: {}".format(mm))

    def early_stop_step(self, metric):
        """:44-62 -- snapshots of theta and every theta_d stay on the device."""
        def snapshot():
            self.best_shared_weights = self.meta_weights.clone()
            self.best_domain_weights = {k: v.clone() for k, v in self.domain_weights.items()}
            self.save_model(self.checkpoint_path)
        if self.base_model.best_metric is None:
            self.base_model.best_metric = metric
            snapshot()
        elif metric <= self.base_model.best_metric:
            self.base_model.counter += 1
            self.log(f'EarlyStopping counter: {self.counter} out of {self.patience}, Best AUC: {self.best_metric}')
            if self.base_model.counter >= self.patience:
                self.base_model.early_stop = True
        else:
            snapshot()
            self.base_model.best_metric = metric
            self.base_model.counter = 0
        return self.base_model.early_stop

    def val_and_test(self, mode):
        """:64-97"""
        if mode == "val":
            dataset = self.dataset.val_dataset
            shared, specific = self.meta_weights, self.domain_weights
        elif mode == "test":
            dataset = self.dataset.test_dataset
            self.load_model(self.checkpoint_path)
            shared, specific = self.best_shared_weights, self.best_domain_weights
        else:
            raise ValueError("Mode can be either val or test, not: {}".format(mode))
        domain_loss, domain_auc = {}, {}
        all_loss, all_auc = 0, 0
        for idx, d in dataset.items():
            self._set_model_merged(shared, specific[idx])
            p_loss, p_auc = self.model.evaluate(d['data'], steps=d['n_step'])
            domain_loss[idx], domain_auc[idx] = p_loss, p_auc
            all_loss += p_loss
            all_auc += p_auc
        avg_loss = all_loss / len(domain_loss)
        avg_auc = all_auc / len(domain_auc)
        self.log("Loss: ", domain_loss)
        self._format_print_domain_metric("AUC", domain_auc)
        weighted_auc = self._weighted_auc(mode, domain_auc)
        self.log("Overall {} Loss: {}, AUC: {}, Weighted AUC: {}".format(mode, avg_loss, avg_auc, weighted_auc))
        return avg_loss, avg_auc, domain_loss, domain_auc

    def _set_model_merged(self, shared, specific):
        """``_set_model_meta_parms(_merge_weights(shared, specific))`` fused: model <- shared (+|*) specific."""
        m = self.model
        for n, (dst, a, b) in self._ranges(m.params, shared.flat, specific.flat):
            m.ctx.call("mamdr_merge", _ptr(dst), _ptr(a), _ptr(b), n, self._merge_method(), m.stream)
            m.ctx.launches += 1

    def _merge_weights(self, shared_weights, specific_weights):
        """:164-172 -- returns a new weight set (theta (+|*) theta_i)."""
        m = self.model
        out = MetaWeights(torch.empty_like(shared_weights.flat), shared_weights.layout, shared_weights.ranges)
        out.flat.copy_(shared_weights.flat)
        for n, (dst, a, b) in self._ranges(out.flat, shared_weights.flat, specific_weights.flat):
            m.ctx.call("mamdr_merge", _ptr(dst), _ptr(a), _ptr(b), n, self._merge_method(), m.stream)
            m.ctx.launches += 1
        return out

    def init_layer(self, model):
        """:174-178 -- re-run every layer's initialiser: a fresh independent draw into the live model."""
        w = self.base_model.draw_initial_weights()
        model.params.copy_(torch.from_numpy(model.layout.pack(w)))

    def separate_train_val_test(self, init_parms=True):
        """:99-162 -- the ``finetune`` stage: per domain, start from best theta (+|*) theta_d, plain SGD
        (lr 0.001), Keras ``EarlyStopping(val_AUC, patience, min_delta=1e-4, mode=max)`` +
        best-``val_AUC`` checkpoint, then test."""
        if init_parms:
            raise NotImplementedError("separate training from scratch is outside the hot path")
        hook = getattr(self, "_discard_lookahead", None)   # a staged look-ahead meta-step will not run
        if hook is not None:
            hook()
        m = self.model
        weights = m.get_weights()                                   # :116 save init weight
        domain_loss, domain_auc = {}, {}
        all_loss, all_auc = 0, 0
        ckpt_dir = osp.dirname(self.checkpoint_path)
        for domain_idx, train_d in self.dataset.train_dataset.items():
            m.compile(optimizer="sgd", lr=0.001)                    # :118-122
            m.set_weights(weights)
            self._set_model_merged(self.best_shared_weights, self.best_domain_weights[domain_idx])
            self.log("Train on domain: {}".format(domain_idx))
            if not osp.exists(ckpt_dir):
                os.makedirs(ckpt_dir)
            val_d = self.dataset.val_dataset[domain_idx]
            best_w = keras_fit_with_callbacks(m, lambda: self.run_train_pass(domain_idx),      # Keras fit(epochs=...) with the
                                              lambda: m.evaluate(val_d['data'], steps=val_d['n_step']),   # two callbacks, :131-142
                                              self.train_config['epoch'], self.train_config['patience'])
            m.set_weights(best_w)                                   # :143 load_weights(chk_path)
            m.save_weights(osp.join(ckpt_dir, "domain_{}{}".format(domain_idx, WEIGHT_EXT)), best_w)
            test_d = self.dataset.test_dataset[domain_idx]
            p_loss, p_auc = m.evaluate(test_d['data'], steps=test_d['n_step'])
            domain_loss[domain_idx], domain_auc[domain_idx] = p_loss, p_auc
            all_loss += p_loss
            all_auc += p_auc
        m.set_weights(weights)                                      # :155 restore
        m.compile(optimizer="adam")
        avg_loss = all_loss / len(domain_loss)
        avg_auc = all_auc / len(domain_auc)
        self.log("Loss: ", domain_loss)
        self._format_print_domain_metric("AUC", domain_auc)
        weighted_auc = self._weighted_auc("test", domain_auc)
        self.log("Overall {} Loss: {}, AUC: {}, Weighted AUC: {}".format("test", avg_loss, avg_auc, weighted_auc))
        return avg_loss, avg_auc, domain_loss, domain_auc
