"""CPU oracle for the MAMDR meta-training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mamdr_b200/`` (the product) may
import this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs execute it, and
there only as the checker or the timed CPU baseline.

What it restates (citations are into ``/root/reference``):

* control flow and meta algebra: ``model_zoo/mamdr.py``,
  ``model_zoo/domain_negotiation.py``,
  ``model_zoo/specific_base_model.py:44-97,164-178``,
  ``model_zoo/maml.py:153-194,343-353``,
  ``model_zoo/base_model.py:111-175,202-224``;
* model topology: ``model_zoo/DeepCTR/deepctr.py:95-136``;
* streaming AUC: ``utils/auc.py:110-157,248-281`` and
  ``utils/metrics_utils.py:245-354``;
* batching contract: ``utils/dataset.py:12-38,73-99``.

PARITY STATUS: the HOST SIDE is pinned to the reference by executing the reference's own code (TF / deepctr stubbed,
``tests/golden/make_reference_golden.py``): the training loops ``MAMDR.train`` / ``DomainNegotiation.train`` /
``Reptile.train`` over a toy Keras stand-in (sequence of train steps, theta, theta_d, best snapshots, early stop: the
oracle's loops replay them bit for bit), and the META ALGEBRA (``mamdr.py:168-196``, ``domain_negotiation.py:118-123``,
``specific_base_model.py:164-172``, ``reptile.py:127-142``) and the early-stop / weighted-AUC
bookkeeping (``base_model.py:157-175,208-224``) are PINNED to the reference: those methods are plain
Python + numpy, ``tests/golden/make_reference_golden.py`` executes the reference's own code (TF / deepctr
stubbed out) and ``tests/test_reference_golden.py`` holds the oracle to the committed vectors bit for bit.
Also pinned by execution: the streaming AUC (``utils/auc.py`` + ``utils/metrics_utils.py``, numpy analogues of the TF
ops) and the forward of the STAR layers (``Star/partitioned_norm.py``, ``Star/star_fcn.py``).
**Parity unpinned** for the rest of the train step (everything below ``model.fit`` / ``train_on_batch``: deepctr's DNN /
MMOE / PLE, Keras BCE, TF Adam, dropout):
The arithmetic of the train step lives in un-vendored third-party packages
(``tensorflow-gpu==1.12.0``, ``deepctr==0.9.0``; ``requirements.txt:1,6``)
which cannot be installed in this image (no wheels for CPython 3.12, no
network).  The oracle restates their published algorithms (SURVEY.md
Appendix A).  The only golden vector the reference holds for this path is the
AUC doc-string example (``utils/auc.py:44-56``); ``tests/test_oracle_auc.py``
pins the oracle to it.  The Philox4x32-10 generator used for dropout masks is
pinned to the Random123 known-answer vectors.
"""
