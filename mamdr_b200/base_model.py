"""``BaseModel`` protocol -- mirrors ``/root/reference/model_zoo/base_model.py`` (paths, early stop,
``val_and_test``, weighted AUC, result files) over the device-resident model of ``engine.py``.
"""
import json
import os
import os.path as osp
import time

from .engine import WEIGHT_EXT
from .schedule import Schedule


def keras_fit_with_callbacks(model, run_pass, evaluate_val, epochs, patience, min_delta=1e-4):
    """``model.fit(..., epochs=epochs, callbacks=[EarlyStopping(monitor='val_AUC', patience, mode='max', min_delta=1e-4),
    ModelCheckpoint(monitor='val_AUC', save_best_only=True, mode='max')])`` as the finetune stage runs it
    (``base_model.py:76-87``, ``specific_base_model.py:131-142``).  [EXT] tf.keras 1.12 callback semantics:
    ModelCheckpoint saves when ``current > best`` (its own running best, no min_delta); EarlyStopping resets its wait only when
    ``current - min_delta > best`` and moves ITS best only then -- a run of sub-``min_delta`` improvements does not ratchet it up;
    it stops when ``wait >= patience``.  Returns the checkpointed weights (the epoch with the best val_AUC)."""
    ck_best, best_w = None, None
    es_best, wait = None, 0
    for _ in range(epochs):
        model.reset_states()
        run_pass()
        _, val_auc = evaluate_val()
        if ck_best is None or val_auc > ck_best:                    # ModelCheckpoint(save_best_only, mode=max)
            ck_best, best_w = val_auc, model.get_weights()
        if es_best is None or val_auc - min_delta > es_best:        # EarlyStopping(min_delta, mode=max): best starts at -inf
            es_best, wait = val_auc, 0
        else:
            wait += 1
            if wait >= patience:
                break
    return best_w if best_w is not None else model.get_weights()


class BaseModel(object):
    def __init__(self, dataset, config):
        self.n_uid = dataset.n_uid
        self.n_pid = dataset.n_pid
        self.n_domain = dataset.n_domain
        self.dataset = dataset
        self.config = config
        self.model_config = config['model']
        self.train_config = config['train']
        self.b200_config = config.get('b200', {})  # optional section the reference ignores

        # base_model.py:23-28 -- same directory layout
        self.checkpoint_path = osp.join(self.train_config['checkpoint_path'], self.model_config['name'],
                                        self.dataset.conf['name'], dataset.conf['domain_split_path'],
                                        time.strftime("%a-%b-%d-%H-%M-%S", time.localtime()),
                                        "model_parameters" + WEIGHT_EXT)
        self.result_path = osp.join(self.train_config['result_save_path'], self.model_config['name'],
                                    dataset.conf['name'], dataset.conf['domain_split_path'])
        from . import dist as mdist
        if mdist.world()[1] > 1:   # the time stamp is taken per process: every rank uses rank 0's path
            import torch.distributed as tdist
            box = [self.checkpoint_path]
            tdist.broadcast_object_list(box, src=0)
            self.checkpoint_path = box[0]
        # injected schedule (python `random` is unseeded in the reference, run.py:26)
        self.schedule = Schedule(self.b200_config.get('schedule_seed', dataset.conf['seed']),
                                 shuffle_batches=not dataset.conf.get('fixed_train', False))   # utils/dataset.py:76
        self.verbose = self.b200_config.get('verbose', True)
        self.model = self.build_model()
        self._build_early_stop()

    def log(self, *a):
        if self.verbose:
            print(*a)

    def build_model(self):
        raise NotImplementedError("You must implement build model")

    def train(self):
        raise NotImplementedError

    # ---- training pass helpers shared by every wrapper ----------------------------------------------
    def stage_epoch_orders(self, passes, mine=None):
        """Draw the sample order of every training pass of one meta-step up-front (in execution order,
        so the draws equal the reference-ordered interleaved ones), write them into ONE pinned host
        buffer and ship them with ONE async H2D copy.  A per-pass upload from pageable memory would
        synchronise the stream before every pass.  ``mine`` (list of bool) selects the passes this rank
        executes when the meta-step is sharded; ids stay global."""
        import collections
        import torch
        first_id = self.schedule.reserve(len(passes))
        ids = [first_id + k for k in range(len(passes))]
        if mine is not None:
            ids = [i for i, m in zip(ids, mine) if m]
            passes = [p for p, m in zip(passes, mine) if m]
        sizes = [self.dataset.train_dataset[i]['n_data'] for i in passes]
        total = int(sum(sizes))
        # two pinned staging buffers used alternately; an event guards the reuse of each
        self._order_flip = 1 - getattr(self, '_order_flip', 1)
        pools = getattr(self, '_order_pools', None)
        if pools is None:
            pools = self._order_pools = [None, None]
        pool = pools[self._order_flip]
        if pool is None or pool[0].numel() < total:
            cap = max(total, 1)
            if pool is not None:
                pool[2].synchronize()
            pool = [torch.empty(cap, dtype=torch.int32).pin_memory(),
                    torch.empty(cap, dtype=torch.int32, device=self.model.device), torch.cuda.Event()]
            pools[self._order_flip] = pool
        else:
            pool[2].synchronize()   # the H2D copy that last read this pinned buffer has completed
        self._order_pool = pool
        host = pool[0].numpy()
        staged, off = collections.deque(), 0
        for i, n, pid in zip(passes, sizes, ids):
            host[off:off + n] = self.schedule.batch_order_at(pid, i, n)
            staged.append((i, off, n))
            off += n
        pool[1][:total].copy_(pool[0][:total], non_blocking=True)
        pool[2].record(torch.cuda.current_stream(self.model.device))
        self._staged_orders = staged
        self.h2d_bytes = getattr(self, 'h2d_bytes', 0) + 4 * total
        return total

    def run_train_pass(self, domain_idx, steps=None):
        """Install the next scheduled sample order of ``domain_idx`` and run one pass (async)."""
        d = self.dataset.train_dataset[domain_idx]
        data = d['data']
        staged = getattr(self, '_staged_orders', None)
        order = None
        if staged:
            i, off, n = staged.popleft()
            assert i == domain_idx and n == data.n_data, "staged schedule out of step with the loop"
            order = self._order_pool[1][off:off + n]   # the pass reads its window of the staged pool directly
        else:
            data.set_order(self.schedule.batch_order(domain_idx, data.n_data))
        n = d['n_step'] if steps is None else steps
        self.samples_trained = getattr(self, 'samples_trained', 0) + min(data.n_data, n * data.batch_size)
        self.last_pass_losses = self.model.fit_pass(data, n, order=order)
        return self.last_pass_losses

    def separate_train_val_test(self, init_parms=True):
        """base_model.py:41-109.  ``init_parms=True`` (`<name>_separate`, run.py:67-68): one model per domain trained from a fresh
        initialisation with the compiled Adam.  ``init_parms=False`` -- the ``finetune`` stage of the wrappers WITHOUT domain-specific
        weights (``*_meta_domain_negotiation_finetune``, ``*_meta_reptile_finetune``; run.py:82-85 loads the best checkpoint
        first): per domain restart from those weights, plain SGD with ``train_config['learning_rate']`` (:69), Keras
        EarlyStopping(val_AUC) + best-val_AUC checkpoint, load it, test; the starting weights are restored at the end."""
        import torch
        hook = getattr(self, "_discard_lookahead", None)   # a staged look-ahead meta-step will not run
        if hook is not None:
            hook()
        m = self.model
        if init_parms:
            # :61-63 `<name>_separate`: tf.global_variables_initializer() -- every variable is re-drawn (a fresh injected draw)
            # and the optimizer slots are zeroed; the compiled Adam is KEPT (no re-compile below) and threads through all domains
            m.params.copy_(torch.from_numpy(m.layout.pack(self.draw_initial_weights())))
            m.reset_optimizer()
        weights = m.get_weights()                                   # :65 save init weight
        domain_loss, domain_auc = {}, {}
        all_loss, all_auc = 0, 0
        ckpt_dir = osp.dirname(self.checkpoint_path)
        for domain_idx, train_d in self.dataset.train_dataset.items():
            if not init_parms:
                m.compile(optimizer="sgd", lr=self.train_config['learning_rate'])    # :67-71
            m.set_weights(weights)                                  # :72
            self.log("Train on domain: {}".format(domain_idx))
            if not osp.exists(ckpt_dir):
                os.makedirs(ckpt_dir)
            val_d = self.dataset.val_dataset[domain_idx]
            best_w = keras_fit_with_callbacks(m, lambda: self.run_train_pass(domain_idx),
                                              lambda: m.evaluate(val_d['data'], steps=val_d['n_step']),
                                              self.train_config['epoch'], self.train_config['patience'])
            m.set_weights(best_w)                                   # :89 load_weights(chk_path)
            m.save_weights(osp.join(ckpt_dir, "domain_{}{}".format(domain_idx, WEIGHT_EXT)), best_w)
            test_d = self.dataset.test_dataset[domain_idx]
            p_loss, p_auc = m.evaluate(test_d['data'], steps=test_d['n_step'])
            domain_loss[domain_idx], domain_auc[domain_idx] = p_loss, p_auc
            all_loss += p_loss
            all_auc += p_auc
        m.set_weights(weights)                                      # :102 restore
        m.compile(optimizer="adam")
        avg_loss = all_loss / len(domain_loss)
        avg_auc = all_auc / len(domain_auc)
        self.log("Loss: ", domain_loss)
        self._format_print_domain_metric("AUC", domain_auc)
        weighted_auc = self._weighted_auc("test", domain_auc)
        self.log("Overall {} Loss: {}, AUC: {}, Weighted AUC: {}".format("test", avg_loss, avg_auc, weighted_auc))
        return avg_loss, avg_auc, domain_loss, domain_auc

    def val_and_test(self, mode):
        """base_model.py:111-144"""
        if mode == "val":
            dataset = self.dataset.val_dataset
        elif mode == "test":
            dataset = self.dataset.test_dataset
            self.load_model(self.checkpoint_path)  # Load best model weights
        else:
            raise ValueError("Mode can be either val or test, not: {}".format(mode))
        domain_loss, domain_auc = {}, {}
        all_loss, all_auc = 0, 0
        for idx, d in dataset.items():
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

    def _format_print_domain_metric(self, name, domain_metric):
        self.log(f"{name}: ")
        for key, value in domain_metric.items():
            self.log(f"{key}: {value}")

    def _weighted_auc(self, mode, domain_auc):
        """base_model.py:157-175"""
        data_info = self.dataset.dataset_info
        tag = 'n_train'
        if "val" in mode:
            tag = "n_val"
        elif "test" in mode:
            tag = "n_test"
        weighted_auc, total_num = 0, 0
        for key, value in domain_auc.items():
            weighted_auc += data_info[key][tag] * value
            total_num += data_info[key][tag]
        return weighted_auc / total_num

    def save_model(self, path):
        from . import dist as mdist
        if mdist.world()[0] != 0:      # replicas are bit-identical: rank 0 writes
            return
        if not osp.exists(osp.dirname(path)):
            os.makedirs(osp.dirname(path))
        self.model.save_weights(path)

    def load_model(self, path):
        from . import dist as mdist
        if mdist.world()[1] > 1:   # rank 0 wrote it: order the read after the write (every rank calls load_model)
            import torch.distributed as tdist
            tdist.barrier()
        self.model.load_weights(path)

    def save_result(self, avg_loss, avg_auc, domain_loss, domain_auc):
        """base_model.py:183-200 -- same files, same names."""
        from . import dist as mdist
        if mdist.world()[0] != 0:
            return None
        result_folder_name = "loss_{:.3f}_auc_{:.3f}_{}".format(avg_loss, avg_auc,
                                                                time.strftime("%a-%b-%d-%H-%M-%S", time.localtime()))
        result_path = osp.join(self.result_path, result_folder_name)
        if not osp.exists(result_path):
            os.makedirs(result_path)
        with open(osp.join(result_path, "dataset_info.json"), 'w') as f:
            json.dump(self.dataset.dataset_info, f)
        with open(osp.join(result_path, "config.json.example"), 'w') as f:
            json.dump(self.config, f)
        with open(osp.join(result_path, "result.json"), 'w') as f:
            json.dump({"avg_loss": avg_loss, "avg_auc": avg_auc, "domain_loss": domain_loss,
                       "domain_auc": domain_auc}, f)
        self.save_model(osp.join(result_path, "model_parameters" + WEIGHT_EXT))
        return result_path

    def _build_early_stop(self):
        self.patience = self.train_config['patience']
        self.counter = 0
        self.best_metric = None
        self.early_stop = False

    def early_stop_step(self, metric):
        """base_model.py:208-224 (strict improvement required)."""
        if self.best_metric is None:
            self.best_metric = metric
            self.save_model(self.checkpoint_path)
        elif metric <= self.best_metric:
            self.counter += 1
            self.log(f'EarlyStopping counter: {self.counter} out of {self.patience}, Best AUC: {self.best_metric}')
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.save_model(self.checkpoint_path)
            self.best_metric = metric
            self.counter = 0
        return self.early_stop
