"""CPU oracle for the PSSR2 test/predict hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pssr2_b200/`` may import this package: it is
the checker, never the thing that is measured or shipped.  Allowed importers are
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.

Each function restates one piece of the reference (``/root/reference``, PSSR2 v2.4.0) or
of a third-party routine the reference calls, and cites the reference ``file:line`` it
follows.  Pinning status (see DESIGN.md "Oracle"):

* Pillow resize restatement  -> pinned against Pillow itself (installed) in tests.
* tiling / crop / pad / stitch / normalize_preds / crappifier arithmetic / ResUNet /
  RDResUNet forward -> pinned against the *reference's own code* imported in the build
  container (``oracle/refshim.py``); outputs committed under ``tests/golden/``.
* scikit-image PSNR / SSIM / ``random_noise`` and timm ``LayerNorm2d`` /
  ``EffectiveSEModule`` -> restated from the published algorithms; those packages are not
  installed anywhere in this environment, so these pieces are **parity unpinned** against
  the third-party code (they are pinned only against closed-form / brute-force checks).
"""
