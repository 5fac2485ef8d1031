"""CPU oracle for the coordinate-network hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement (torch fp32 / int64 on CPU, numpy for byte
work) of the algorithms that Benjamin-Fouquet/mri_interpolation runs on its hot
path.  Every function cites the reference file:line it follows.

Rules (see DESIGN.md "Oracle"):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * the product package ``mri_interpolation_b200`` never imports it and has no
    CPU fallback: it raises if the CUDA library is missing;
  * parity is PINNED: ``oracle/make_golden.py`` imports the reference's own
    ``encoding.py`` / ``models.py`` from ``/root/reference`` (with stub modules
    for the absent third-party packages), asserts that this restatement is
    bit-identical to it on CPU, and writes the fixtures under ``tests/golden/``.
    SSIM (skimage, absent) is the one item that stays "parity unpinned".
"""
