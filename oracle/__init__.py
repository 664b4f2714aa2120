"""CPU oracle for the CACC / A2C / NeurComm hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it, and there only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product path
(``deeprl_network_b200``) never imports this package and fails loudly when the
CUDA extension is missing.

Pinning status
--------------
* ``oracle.cacc``     (env)      -- PINNED: checked bit-for-bit against trajectories
  produced by importing the unmodified reference ``envs/cacc_env.py``
  (``tests/golden/make_golden.py`` -> ``tests/golden/env_*.npz``) and against the
  known-answer values in SURVEY.md section 8(c).
* ``oracle.buffers``  (returns)  -- PINNED: checked against the unmodified reference
  ``agents/utils.py`` buffers (``tests/golden/buffer_*.npz``).
* ``oracle.nets`` / ``oracle.trainer`` (TF1 graphs, loss, optimizer, loop) --
  PARITY UNPINNED: TensorFlow 1.12 is not installable here and the reference has no
  tests at this boundary, so these are a line-by-line restatement of
  ``agents/utils.py`` / ``agents/policies.py`` / ``agents/models.py`` / ``utils.py``
  in PyTorch-CPU (fp32, fp64 switch) reviewed against the cited lines only.
"""
