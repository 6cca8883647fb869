"""``DeepCTR`` base model (mlp tower) -- mirrors ``/root/reference/model_zoo/DeepCTR/deepctr.py``:
``build_inputs`` / ``build_emb`` (:95-116), ``build_mlp`` (:118-136), compile (:54-60) and the joint
'alternate' ``train`` loop (:63-93).  Only the ``mlp`` tower is on the hot path.
"""
import time

from . import _lib
from .base_model import BaseModel
from .engine import MLPModel
from .layout import init_mlp_weights, mlp_layout

_PRECISION = {"fp32": _lib.PREC_FP32, "tf32": _lib.PREC_TF32, "tf32x3": _lib.PREC_TF32X3}


class DeepCTR(BaseModel):
    def __init__(self, dataset, config):
        super(DeepCTR, self).__init__(dataset, config)

    def build_model(self):
        name = self.model_config['name']
        if 'mlp' not in name:
            raise NotImplementedError(
                "tower {!r}: only 'mlp' is on the B200 hot path (wdl/nfm/autoint/ccpm/pnn/deepfm are out of scope)"
                .format(name))
        mc, tc = self.model_config, self.train_config
        # deepctr.py:105-116 -- emb_trainable is honoured only with load_pretrain_emb; otherwise the
        # tables are default-initialised and always trainable
        if tc['load_pretrain_emb']:
            if self.dataset.user_table is None or self.dataset.item_table is None:
                raise AttributeError("dataset has no pretrained user_emb / item_emb")
            emb_trainable = bool(tc['emb_trainable'])
        else:
            emb_trainable = True
        self.emb_trainable = emb_trainable
        emb_dim = (mc['user_dim'], mc['item_dim'], mc['domain_dim'])
        self.layout = mlp_layout(self.n_uid, self.n_pid, self.n_domain, emb_dim, mc['hidden_dim'], emb_trainable)
        self._init_draws = 0
        self.init_seed = self.b200_config.get('init_seed', self.dataset.conf['seed'])
        w0 = self.draw_initial_weights()
        if emb_trainable and tc['load_pretrain_emb']:
            w0[self.layout.index('user_emb')] = self.dataset.user_table
            w0[self.layout.index('item_emb')] = self.dataset.item_table
        if tc['optimizer'] != 'adam':
            raise NotImplementedError("only the 'adam' optimizer is on the hot path")
        if tc['loss'] != 'binary_crossentropy':
            raise NotImplementedError("only binary_crossentropy is on the hot path")
        model = MLPModel(self.n_uid, self.n_pid, self.n_domain, emb_dim=emb_dim, hidden=tuple(mc['hidden_dim']),
                         dropout=mc.get('dropout', 0.0), dropout_seed=1024, l2_emb=1e-5,
                         emb_trainable=emb_trainable,
                         user_table=None if emb_trainable else self.dataset.user_table,
                         item_table=None if emb_trainable else self.dataset.item_table,
                         init_weights=w0, lr=tc['learning_rate'], max_batch=self.dataset.batch_size,
                         precision=_PRECISION[self.b200_config.get('precision', 'fp32')],
                         device=self.b200_config.get('device', self.dataset.device),
                         use_graphs=self.b200_config.get('cuda_graphs', True))
        return model

    def draw_initial_weights(self):
        """k-th independent draw of every layer's initialiser (k = 0 is the model build; k >= 1 are the
        ``init_layer`` re-initialisations of specific_base_model.py:174-178)."""
        w = init_mlp_weights(self.layout, [self.init_seed, self._init_draws])
        self._init_draws += 1
        return w

    def train(self):
        """deepctr.py:63-93 -- joint training: shuffled domains, one full pass each, one Adam."""
        self.model.reset_optimizer()
        train_sequence = list(range(self.n_domain))
        for epoch in range(self.train_config['epoch']):
            self.log("Epoch: {}".format(epoch), "-" * 30)
            train_sequence = self.schedule.shuffle_sequence(train_sequence)
            self.stage_epoch_orders(list(train_sequence))
            for idx in train_sequence:
                self.log("Train on: Domain {}".format(idx))
                old_time = time.time()
                self.model.reset_states()
                self.run_train_pass(idx)
                self.log("Training time: ", time.time() - old_time)
            self.log("Val Result: ")
            avg_loss, avg_auc, domain_loss, domain_auc = self.val_and_test("val")
            if self.early_stop_step(avg_auc):
                break
            self.log("Test Result: ")
            self.val_and_test("test")
