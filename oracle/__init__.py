r"""CPU oracle for the SDA hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker (or as the timed CPU
baseline), never as the thing shipped.  The product path (``sda_b200``) fails
loudly when its CUDA library is missing; it never routes through this package.

Contents
--------
``score_oracle``      plain torch-CPU (fp32 / fp64) restatement of the score path:
                      window maps, U-Net, time embedding, VPSDE schedule, sampler,
                      Gaussian guidance.  Every function cites the reference
                      file:line it restates.  PINNED: checked against the
                      reference classes imported from ``/root/reference`` (with the
                      zuko shim of ``ref_import``) by ``tests/golden/make_golden.py``;
                      the resulting vectors are committed under ``tests/golden/``.
``kolmogorov_oracle`` NumPy restatement of the jax-cfd finite-volume scheme that
                      ``sda.mcs.KolmogorovFlow`` drives.  PARITY UNPINNED: jax /
                      jax-cfd are absent from the reference tree and from this
                      image, and the reference holds no test or golden vector for
                      it.
``ref_import``        helper that imports the unmodified reference package from
                      ``/root/reference`` (only exists in the build container).
"""
