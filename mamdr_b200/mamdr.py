"""MAMDR = Domain Negotiation on the shared parameters + Domain Regularization on the
domain-specific parameters -- mirrors ``/root/reference/model_zoo/mamdr.py`` (``train`` :18-166 and
the update helpers :168-196).  Control flow is the reference's; every weight set is a device arena
and every update is one fused sweep in ``libmamdr_b200.so`` (no host round trips).
"""
import torch

from . import dist as mdist
from .engine import _ptr
from .maml import MetaWeights
from .specific_base_model import SpecificBase


class MAMDR(SpecificBase):
    def __init__(self, base_model):
        super(MAMDR, self).__init__(base_model)

    def train(self):
        self.log("Start MAMDR on model: {}".format(self.model_config['name']))
        self.prepare()
        for epoch in range(self.train_config['epoch']):              # :41
            self.log("Epoch: {}".format(epoch), "-" * 30)
            self.train_epoch(epoch)
            if epoch % self.train_config['val_every_step'] == 0:     # :145-159
                val_avg_loss, val_avg_auc, val_domain_loss, val_domain_auc = self.val()
                val_metric = val_domain_auc[self.train_config['target_domain']] \
                    if self.train_config['target_domain'] >= 0 else val_avg_auc
                if self.early_stop_step(val_metric):
                    break
                self.log("Test Result: ")
                self.val_and_test("test")
        self._discard_plan()

    def _discard_plan(self):
        """Drop a look-ahead plan that will not run (early stop, last epoch, the finetune stage follows): the schedule goes back
        to its state from before the look-ahead, so whatever draws next sees exactly what the reference order implies."""
        nxt = getattr(self, "_next_plan", None)
        self._next_plan = None
        if nxt is not None and nxt["schedule"] is self.schedule:
            self.schedule._rng.setstate(nxt["rng_before"][0])
            self.schedule._pass = nxt["rng_before"][1]
            self.base_model._staged_orders = None

    def prepare(self):
        """:26-37 -- theta = the model's first initialisation; theta_d^0 = an independent
        re-initialisation per domain; optimizer slots zeroed; meta sequence built."""
        self._get_model_meta_parms()
        self.meta_weights = self._get_meta_weights()
        self.domain_weights = {}
        for domain_idx in range(self.n_domain):
            self.init_layer(self.model)
            self.domain_weights[domain_idx] = self._get_meta_weights()
        self.model.reset_optimizer()
        self.train_sequence = self.build_meta_data_split()
        self._accum = None
        self._setup_lanes()

    # ---- "virtual ranks" (opt-in, b200.virtual_ranks = V > 1, one process / one GPU): the DR chains of different query domains
    # run CONCURRENTLY on V models ("lanes"), each lane's persistent pass kernel on 1 / V of the SMs and on its own stream.  The
    # semantics are exactly those of the V-rank sharded schedule (replicated DN -> here: DN once at the full grid, then copied;
    # LPT-assigned chains with per-rank Adam state; the last owner's optimizer state and live model adopted by all), judged
    # against `OracleMAMDR.train_epoch_sharded(V)`.  A row-local chain keeps 64 of the 148 SMs busy: two
    # chains side by side fill the machine (SURVEY.md 7.3 hard part 1(c)).  Never the default: V = 1 is the reference schedule.
    def _setup_lanes(self):
        self._lane_models = None
        V = int(self.b200_config.get('virtual_ranks', 1))
        if V <= 1:
            return
        if mdist.world()[1] > 1:
            raise NotImplementedError("b200.virtual_ranks applies to single-process runs (use torchrun ranks OR virtual ranks)")
        m = self.model
        if not getattr(m, "pass_kernel", False) or m.emb_trainable or not hasattr(m, "clone_lane"):
            raise NotImplementedError("b200.virtual_ranks needs the persistent pass kernel (frozen-table mlp, tf32 / tf32x3)")
        if self.train_config['finetune_every_epoch'] or "batch" in self.model_config['name']:
            raise NotImplementedError("b200.virtual_ranks with finetune_every_epoch / 'batch' names (their sharded schedule shards pairs)")
        dev = m.device
        self._lane_models = [(m, None)] + [(m.clone_lane(), torch.cuda.Stream(device=dev)) for _ in range(V - 1)]
        self._lane_ctas = max(32, m.ctx.sm_count // V)
        for lm, _ in self._lane_models[1:]:
            lm.set_pass_ctas(self._lane_ctas)
        self._lane_accum = [None] * V

    def _n_shards(self):
        world = mdist.world()[1]
        if world > 1:
            return world
        return len(self._lane_models) if getattr(self, "_lane_models", None) else 1

    def _plan_epoch(self):
        """Everything of a meta-step that does not depend on the weights: the schedule draws (sequence shuffle :45-46, DR support
        samples :66-70, in the reference's order), the multi-GPU assignment, and the staging of every pass's sample order (host
        permutations + ONE pinned H2D copy).  Run for meta-step k+1 right after the launches of meta-step k were enqueued, it
        overlaps ~10 ms of host work with the GPU instead of serialising it behind the step's result read-back."""
        tc = self.train_config
        sched = self.schedule
        plan = {"schedule": sched, "rng_before": (sched._rng.getstate(), sched._pass), "sequence_before": list(self.train_sequence)}
        train_sequence = self.train_sequence
        if tc['shuffle_sequence']:                                   # :45-46
            train_sequence = sched.shuffle_sequence(train_sequence)
        # the DR support samples (:66-70) are drawn up-front, in the reference's order (the per-pass
        # sample orders come from an independent keyed stream), so the whole meta-step can be staged
        supports = {}
        for idx in train_sequence:
            candidate_domains = list(train_sequence)
            candidate_domains.remove(idx)
            aux_idxs = sched.sample_support(candidate_domains, tc['sample_num'])   # :68
            if tc['add_query_domain']:
                aux_idxs = list(aux_idxs) + [idx]
            supports[idx] = aux_idxs
        # multi-GPU: DN replicated, DR query domains LPT-sharded (mamdr_b200/dist.py)
        rank, world = mdist.world()
        n_step = {i: self.dataset.train_dataset[i]['n_step'] for i in train_sequence}
        batch_mode = "batch" in self.model_config['name']
        # 'batch' names (:100-108): every (query i, support j) pair starts from the same theta (+|*) theta_i, so the PAIRS are
        # the independent units (60 at Taobao-10 instead of 10 chains): LPT-sharded by S_j + S_i, the accumulated deltas of
        # all theta_i all-reduced once per meta-step.  finetune_every_epoch chains a pass per domain behind the update:
        # those configs keep the chain sharding.
        pair_mode = batch_mode and world > 1 and not tc['finetune_every_epoch']
        owner = mdist.lpt_assign(mdist.dr_chain_costs(train_sequence, supports, n_step,
                                                      tc['domain_regulation_step']), self._n_shards())
        pair_owner = None
        if pair_mode:
            pair_owner = mdist.lpt_assign(mdist.dr_pair_costs(train_sequence, supports, n_step,
                                                              tc['domain_regulation_step']), world)
            owner = {idx: pair_owner[(pos, len(supports[idx]) - 1)] for pos, idx in enumerate(train_sequence)}
        passes, mine = list(train_sequence), [True] * len(train_sequence)
        for pos, idx in enumerate(train_sequence):
            for k, aux_idx in enumerate(supports[idx]):
                passes += [aux_idx, idx]
                mine += [(pair_owner[(pos, k)] if pair_mode else owner[idx]) == rank] * 2
            if tc['finetune_every_epoch']:
                passes.append(idx)
                mine.append(owner[idx] == rank)
        self.stage_epoch_orders(passes, mine if world > 1 else None)
        plan.update(train_sequence=train_sequence, supports=supports, owner=owner, pair_owner=pair_owner, pair_mode=pair_mode,
                    batch_mode=batch_mode, rank=rank, world=world,
                    staged=(getattr(self.base_model, "_staged_orders", None), getattr(self.base_model, "_order_pool", None)))
        return plan

    def train_epoch(self, epoch=0):
        """One MAMDR meta-step: the body of the epoch loop, :44-143."""
        tc = self.train_config
        beta = tc['meta_learning_rate']
        plan = getattr(self, "_next_plan", None)
        self._next_plan = None
        if plan is None or plan["schedule"] is not self.schedule or plan["sequence_before"] != list(self.train_sequence):
            plan = self._plan_epoch()
        elif plan["staged"][0] is not None:      # the staged orders of the prefetched plan become the live ones
            self.base_model._staged_orders, self.base_model._order_pool = plan["staged"]
        self.train_sequence = train_sequence = plan["train_sequence"]
        supports, owner, pair_owner, pair_mode, batch_mode = plan["supports"], plan["owner"], plan["pair_owner"], plan["pair_mode"], plan["batch_mode"]
        rank, world = plan["rank"], plan["world"]
        self.dr_owner, self.dr_pair_owner = owner, pair_owner

        # In the tcgen05 modes the DN phase and every DR chain of this rank are each recorded and run as ONE persistent
        # launch (engine.program: passes + meta sweeps executed in-kernel); recording chain k+1 overlaps the execution
        # of chain k.  finetune_every_epoch allocates temporaries -> immediate mode.
        use_program = self.b200_config.get('program', True) and not tc['finetune_every_epoch']
        with self.model.program(use_program):
            # ---- Update Shared (DN), :48-57
            self._set_model_meta_parms(self.meta_weights)
            for idx in train_sequence:
                self.run_train_pass(idx)
            self._update_meta_weight(self.meta_weights, meta_lr=beta)

        # ---- Update specific (DR), :59-108
        if pair_mode:
            self._dr_pairs_sharded(train_sequence, supports, pair_owner, rank, use_program)
            self._prefetch_plan()
            return
        lanes = getattr(self, "_lane_models", None) if world == 1 else None
        if lanes:
            self._dr_chains_on_lanes(train_sequence, supports, owner, batch_mode, beta, use_program, lanes)
            self._prefetch_plan()
            return
        for idx in train_sequence:
            if owner[idx] != rank:
                continue
            self._dr_chain(idx, supports[idx], batch_mode, beta, use_program)

        if world > 1:
            # the one collective of the meta-step: theta_i from their owners + the Adam slots of the rank
            # that owns the last query domain of the sequence
            m = self.model
            words = m.opt_words()
            # the live model of the last chain's owner too: with a subset of meta parameters (config #4: STAR) the other
            # variables and the PartitionedNorm moving statistics train through every pass and are never reloaded from theta
            extra = [m.params]
            pn_steps = None
            if getattr(m, "pn_state", None) is not None:
                pn_f, pn_steps = m.pn_parts()          # float statistics; int32 update counters (exact as fp32 below 2^24)
                steps_f = pn_steps.to(torch.float32)
                extra += [pn_f, steps_f]
            self.comm_bytes = mdist.exchange(owner, rank, {k: v.flat for k, v in self.domain_weights.items()},
                                             m.m, m.v, words, owner[train_sequence[-1]], extra)
            if pn_steps is not None:
                pn_steps.copy_(steps_f.to(torch.int32))
            m.set_opt_words(words)
        self._prefetch_plan()

    def _dr_chain(self, idx, aux_idxs, batch_mode, beta, use_program):
        """The DR chain of query domain `idx` (:62-143) on the current model / stream."""
        tc = self.train_config
        with self.model.program(use_program):
            d = self.dataset.train_dataset[idx]
            theta_i = self.domain_weights[idx]
            # merged = theta (+|*) theta_i is never materialised on the host: model <- merged (:72,78)
            self._set_model_merged(self.meta_weights, theta_i)
            if batch_mode:
                self._zero_accum()
            for k, aux_idx in enumerate(aux_idxs):
                self.log(f"Support Domain: {aux_idx}, Query Domain: {idx}")
                self.run_train_pass(aux_idx)                         # :85-86
                train_step = d['n_step']                             # :92-97
                if tc['domain_regulation_step'] > 0:
                    train_step = min(train_step, tc['domain_regulation_step'])
                self.run_train_pass(idx, train_step)
                if batch_mode:                                       # :100-101
                    self._accumulate_grad(theta_i)
                    self._set_model_merged(self.meta_weights, theta_i)
                else:                                                # :103-105 + next iteration's :78
                    self._dr_update(theta_i, beta)
            if batch_mode:                                           # :107-108
                self._update_meta_weight_by_grads(theta_i)

            if tc['finetune_every_epoch']:                           # :110-143
                merged = self._merge_weights(self.meta_weights, theta_i)
                self._set_model_meta_parms(merged)
                for m in self.model.stateful_metric_functions:
                    m.reset_states()
                self.run_train_pass(idx)
                self._update_domain_weights(theta_i, merged)

    def _dr_chains_on_lanes(self, train_sequence, supports, owner, batch_mode, beta, use_program, lanes):
        """Virtual ranks: chain `idx` runs on lane owner[idx]; lanes differ in model, context and stream, so consecutive
        chains of different lanes execute side by side.  The chains are ENQUEUED in sequence order (the staged sample orders
        are consumed in that order); each lane executes its own chains in sequence order, like a rank of the sharded schedule."""
        base = self.base_model
        lane0 = lanes[0][0]
        dev = lane0.device
        main = torch.cuda.current_stream(dev)
        after_dn = torch.cuda.Event()
        after_dn.record(main)
        for lm, ls in lanes[1:]:
            with torch.cuda.stream(ls):
                ls.wait_event(after_dn)          # theta, the staged orders and lane 0's post-DN state are ready
                lm.copy_state_from(lane0)
        lane0.set_pass_ctas(self._lane_ctas)
        main_accum = self._accum
        try:
            for idx in train_sequence:
                r = owner[idx]
                lm, ls = lanes[r]
                base.model = lm
                self._accum = self._lane_accum[r]
                with torch.cuda.stream(ls if ls is not None else main):
                    self._dr_chain(idx, supports[idx], batch_mode, beta, use_program)
                self._lane_accum[r] = self._accum
        finally:
            base.model = lane0
            self._accum = main_accum
            lane0.set_pass_ctas(0)
        for lm, ls in lanes[1:]:
            e = torch.cuda.Event()
            e.record(ls)
            main.wait_event(e)
        last = owner[train_sequence[-1]]
        if last != 0:                            # the Adam state and live model of the last chain's owner are adopted
            lane0.copy_state_from(lanes[last][0])

    def _prefetch_plan(self):
        """Stage the NEXT meta-step while the GPU still runs this one (`b200.lookahead`, default on).  The draws happen in the
        reference's order either way; `save_state` stores the schedule state from before the look-ahead."""
        if (self.b200_config.get('lookahead', True) and not self.train_config['finetune_every_epoch']
                and not self.train_config.get('meta_finetune_step', 0) > 0):   # (val() would train between the meta-steps)
            self._next_plan = self._plan_epoch()
            self.base_model._discard_lookahead = self._discard_plan

    def _dr_pairs_sharded(self, train_sequence, supports, pair_owner, rank, use_program):
        """DR of a 'batch' name on world > 1 ranks: this rank runs its (query, support) pairs, accumulating the deltas per query
        domain (:182-191); ONE all-reduce sums the accumulators of all theta_i (and carries the Adam slots / live model of the
        rank that owns the last pair of the sequence); every rank then applies :193-196 to every theta_i -- replicas stay
        bit-identical."""
        tc = self.train_config
        m = self.model
        P_ = m.params.numel()
        if getattr(self, "_accum_all", None) is None or self._accum_all.shape[0] != len(train_sequence):
            self._accum_all = torch.zeros(len(train_sequence), P_, dtype=torch.float32, device=m.params.device)
        else:
            self._accum_all.zero_()
        for pos, idx in enumerate(train_sequence):
            ks = [k for k in range(len(supports[idx])) if pair_owner[(pos, k)] == rank]
            if not ks:
                continue
            with self.model.program(use_program):
                d = self.dataset.train_dataset[idx]
                theta_i = self.domain_weights[idx]
                self._accum = self._accum_all[pos]
                self._set_model_merged(self.meta_weights, theta_i)
                for k in ks:
                    aux_idx = supports[idx][k]
                    self.log(f"Support Domain: {aux_idx}, Query Domain: {idx}")
                    self.run_train_pass(aux_idx)   # staged orders are consumed in this (global) order
                    train_step = d['n_step']
                    if tc['domain_regulation_step'] > 0:
                        train_step = min(train_step, tc['domain_regulation_step'])
                    self.run_train_pass(idx, train_step)
                    self._accumulate_grad(theta_i)
                    self._set_model_merged(self.meta_weights, theta_i)
        self._accum = None
        words = m.opt_words()
        last_pos = len(train_sequence) - 1
        last_owner = pair_owner[(last_pos, len(supports[train_sequence[-1]]) - 1)]
        extra = [m.params]
        self.comm_bytes = mdist.exchange_sum(self._accum_all, m.m, m.v, words, rank, last_owner, extra)
        m.set_opt_words(words)
        for pos, idx in enumerate(train_sequence):
            self._accum = self._accum_all[pos]
            self._update_meta_weight_by_grads(self.domain_weights[idx])
        self._accum = None

    # ---- :168-171
    def _update_domain_weights(self, domain_weights, merged_weights):
        m = self.model
        for n, (out, a, b) in self._ranges(domain_weights.flat, m.params, merged_weights.flat):
            m.ctx.call("mamdr_sub", _ptr(out), _ptr(a), _ptr(b), n, m.stream)
            m.ctx.launches += 1

    # ---- :173-180
    def _update_meta_weight(self, update_vars, merged_weights=None, meta_lr=1):
        """update += (model - old) * meta_lr with old = merged_weights if given else update itself."""
        m = self.model
        old = merged_weights if merged_weights is not None else update_vars
        for n, (upd, model, o) in self._ranges(update_vars.flat, m.params, old.flat):
            m.ctx.call("mamdr_axpy_diff", _ptr(upd), _ptr(model), _ptr(o), meta_lr, n, m.stream)
            m.ctx.launches += 1

    def _dr_update(self, theta_i, beta):
        """Fused :103-105 and the following :78: theta_i += (model - (theta (+|*) theta_i)) * beta, then
        model <- theta (+|*) theta_i.  20 B / parameter, one sweep."""
        m = self.model
        for n, (ti, th, model) in self._ranges(theta_i.flat, self.meta_weights.flat, m.params):
            m.ctx.call("mamdr_dr_update", _ptr(ti), _ptr(th), _ptr(model), beta, n, self._merge_method(),
                       _ptr(model), m.stream)
            m.ctx.launches += 1

    # ---- :182-196 ("batch" names)
    def _zero_accum(self):
        if self._accum is None:
            self._accum = torch.zeros_like(self.model.params)
        else:
            self._accum.zero_()

    def _accumulate_grad(self, theta_i):
        m = self.model
        for n, (acc, model, th, ti) in self._ranges(self._accum, m.params, self.meta_weights.flat, theta_i.flat):
            m.ctx.call("mamdr_dr_accumulate", _ptr(acc), _ptr(model), _ptr(th), _ptr(ti), n, self._merge_method(),
                       m.stream)
            m.ctx.launches += 1

    def _update_meta_weight_by_grads(self, theta_i):
        m = self.model
        tc = self.train_config
        for n, (ti, acc) in self._ranges(theta_i.flat, self._accum):
            m.ctx.call("mamdr_dr_apply_accum", _ptr(ti), _ptr(acc), float(tc['sample_num']),
                       tc['meta_learning_rate'], n, m.stream)
            m.ctx.launches += 1
