"""CPU oracle for the InferBiomechanics hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or as the timed CPU baseline.  The
product package (``inferbiomechanics_b200``) never imports it and has no CPU fallback.

Contents
--------
``loss.py``      restatement of ``src/loss/RegressionLossEvaluator.py`` (E-1 … E-6)
``windows.py``   restatement of the window index / sampler / packing rules
                 (``src/data/AddBiomechanicsDataset.py:121-285``, ``src/cli/train.py:143-150``)
``models.py``    functional fp32/fp64 restatement of the three reference models' forward passes
``ddpm.py``      builder-owned DDPM spec (NOT in the reference — *parity unpinned*)
``train.py``     one reference-shaped training step on CPU (forward, loss, backward, RMSprop)
``refimport.py`` imports the real reference from ``/root/reference`` (this container only)
``gen_golden.py`` writes ``tests/golden/*.npz`` from the real reference

Pinning status
--------------
* loss helpers: pinned by the reference's own 24 unit tests (re-stated in
  ``tests/test_oracle_loss.py``) and by golden vectors generated from the imported reference.
* loss ``__call__`` composition, model forwards/backwards: pinned by golden vectors generated
  by importing the reference modules here (``gen_golden.py``).
* window enumeration / packing: the reference class needs nimblephysics + ``.b3d`` files, neither
  of which exist here; the restatement follows the cited lines and is pinned only by
  hand-computed cases — stated as such.
* DDPM (q_sample, timestep embedding, posterior step, denoiser): the reference contains no
  diffusion code at all (only ``src/.gitignore:10``).  **Parity unpinned — builder oracle.**
"""
