"""CPU oracle for the NConv depth-completion hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in plain torch CPU ops (dtype-generic: fp32 or fp64), the algorithm
of the reference's hot path (``model_zoo/nconv.py``, ``model_zoo/nconv_sd.py``,
``model_zoo/confidence_depth_framework.py``, ``model_zoo/s2d_u_resnet.py``,
``model_zoo/s2d_resnet.py``, ``metrics.py``, ``dataloaders/dense_to_sparse.py``
under /root/reference).  Every function cites the reference ``file:line`` it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it -
and only as the checker / the timed CPU baseline, never as something the
product path routes through.  The product package
(``aerial_depth_completion_b200``) must not import ``oracle``.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the *unmodified
reference executed in the build container* (torch 2.11 CPU).  The generating
script is ``tests/golden/make_golden.py``; the fixtures it wrote are committed
under ``tests/golden/*.npz`` and checked by ``tests/test_oracle_golden.py``.
When ``/root/reference`` is present the same tests additionally compare the
oracle with the live reference.
"""

from . import nconv_oracle, nets_oracle, loss_metrics_oracle, sampling_oracle, synth  # noqa: F401
