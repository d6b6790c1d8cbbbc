"""CPU oracle for the DINCAE train + reconstruct hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker (or as the timed CPU
baseline) -- never as a fallback for the CUDA path in ``dincae_b200``.

Contents
--------
``ref_loader``   imports the *unmodified* reference (``/root/reference/DINCAE``) with
                 ``tensorflow`` / ``netCDF4`` stubbed so its numpy input pipeline
                 (``data_generator_list``, closure ``datagen``) runs.  Only usable in the
                 build container (the GPU box has no ``/root/reference``); used by
                 ``oracle/make_golden.py`` to write ``tests/golden/*.npz``.
``datagen_np``   numpy restatement of that input pipeline (DINCAE/__init__.py:120-230).
                 PINNED: bit-exact against the reference generator's own output, see
                 ``tests/golden/datagen_*.npz`` and ``tests/test_oracle_datagen.py``.
``model_torch``  PyTorch-CPU restatement of the TF-1.15 graph built by ``reconstruct``
                 (DINCAE/__init__.py:433-603): encoder, dense bottleneck, decoder, output
                 head, masked Gaussian NLL, global-norm clip, TF-Adam.
                 PARITY UNPINNED at tensor level: the arithmetic lives in
                 ``tensorflow==1.15.4`` (setup.py:19), which is neither vendored in the
                 reference tree nor installable here (no py3.12 wheel), and the reference's
                 tests hold only scalar loss bounds on a file that must be downloaded.  The
                 restatement follows the published TF-1.15 op semantics named in SURVEY.md
                 section 8c and is anchored on the reference's call sites.
"""
