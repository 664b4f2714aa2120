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
* ``oracle.trainer`` (Trainer / Counter control flow, agent-class host logic) -- PINNED: replays,
  bit for bit, traces recorded from the unmodified reference ``utils.Trainer`` and ``agents/models.py``
  classes driving scripted agents / scripted policies (``tests/golden/trainer_*.npz``,
  ``tests/golden/agent_*.npz``; ``tests/test_trainer_flow.py``, ``tests/test_agent_flow.py``).
* ``oracle.nets`` (TF1 graphs, loss, autodiff, clip, optimizer) -- PINNED TO THE REFERENCE SOURCE RUN ON A
  TF SHIM: TensorFlow 1.12 is not installable here, so ``tests/golden/make_golden.py`` executed the
  unmodified reference end to end (env, Trainer, agent, policy and layer code) with ``tensorflow`` replaced
  by ``tests/golden/tf_shim.py`` -- the ~35 TF primitives those files call, restated on PyTorch-CPU -- and
  stored initial weights, every pi / v / R and the trained weights for all six agents
  (``tests/golden/tfnet_*.npz``).  ``tests/test_tfnet_parity.py``: identical initial weights and sampled
  actions, pi / v / R within 2e-7, weights after 4-8 updates within 3e-8.  What remains UNPINNED is only the
  shim's restatement of the TF primitives themselves (array ops, matmul, activations, ``tf.gradients`` via
  autograd, and the ``clip_by_global_norm`` / ``RMSPropOptimizer`` update formulas).
"""
