"""CPU oracle for the dmdqn agent-side hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy for byte/integer/fp64 work, CPU
PyTorch fp32 for the MLP arithmetic) of what the reference
(pranshu-raj-211/dmdqn, TensorFlow/Keras) computes on the path

    featurise -> epsilon-greedy act -> replay push/sample -> Double-DQN learn

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or the timed CPU baseline -- never as (or underneath) the product
path in ``dmdqn_b200/``.

Pinning status (see DESIGN.md "Oracle"):
  * featurisation, reward, replay add/sample/z-score, epsilon schedule are
    PINNED against the reference's own Python code executed in the build
    container with stubbed third-party modules (``oracle/make_golden.py`` ->
    ``tests/golden/ref_*.npz``).
  * the MLP / Double-DQN / Adam arithmetic lives in TensorFlow 2.19 +
    Keras 3.9.2, which are not installable here: that part is restated from
    the published Keras algorithms and cross-checked against
    ``torch.autograd`` and ``torch.optim.Adam`` -- "parity unpinned" by the
    reference for those functions.
"""
