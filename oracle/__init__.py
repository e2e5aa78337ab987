"""CPU oracle for the dino_pose hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker (or as
the timed CPU baseline).  Nothing under ``dino_pose_b200/`` imports it: the product
path is CUDA-only and raises when ``libdinopose_sm100a.so`` is missing.

Contents
--------
* ``weights.py``        deterministic synthetic ``state_dict`` generator keyed by
                        parameter NAME (so the reference, the oracle and the CUDA
                        path can all be loaded with bit-identical weights without
                        shipping 120 MB fixtures).
* ``pose_oracle.py``    plain-torch fp32 functional restatement of the reference
                        forward (HF ``Dinov2Model`` 5.5.0 arithmetic + the reference's
                        LoRA / pose heads), the losses of ``train.py`` and, through
                        autograd, the gradients.
* ``decode_oracle.py``  numpy restatement of the heat-map -> key-point decode.
* ``ref_harness.py`` /
  ``make_golden.py``    import the REAL reference from ``/root/reference`` (build
                        container only) and freeze its outputs into
                        ``tests/golden/*.npz``.  The oracle is pinned against those
                        vectors by ``tests/test_oracle_golden.py`` (CPU, ``-m "not gpu"``).

Parity status: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the pin is "outputs of the reference itself run here"
(``make_golden.py``), committed as fixtures together with the generating script.
"""
