"""CPU oracle for the scoring-and-selection hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a CPU restatement (numpy / a little C) of what the upstream
reference computes on this path.  It exists so that the CUDA path can be checked for parity on a
machine that has no copy of the reference (the GPU box).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker / the timed CPU arm -- never as a fallback for the product
package ``multi_view_active_learning_b200`` (which raises if its CUDA library is missing).

Pinning: the reference's own tests assert shapes only, so the oracle is
pinned against outputs of the *unmodified reference executed in the dev container*
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; the known-answer values of
``tests/test_triangulation.py`` and ``tests/test_coreset.py`` are part of those fixtures).
Functions for which the reference's third-party dependency is not installed anywhere here
(kornia soft-argmax, skimage peak_local_max) are restated from the published algorithm and say
"parity unpinned" in their docstring.
"""
