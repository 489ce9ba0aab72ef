"""CPU oracle for the leaffliction preprocessing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``leaffliction_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the reported CPU baseline, never as the thing shipped.

Two layers live here:

* ``spec_*.py`` -- NumPy restatements ("spec functions") of the arithmetic the
  reference delegates to Pillow 12.2.0 / OpenCV 4.13.0 / NumPy 2.3.5 (the
  versions in this image; the reference pins none, requirements.txt:1-13).
  Each function cites the reference call site (file:line under
  /root/reference) whose result it restates.
* ``refcalls.py`` -- the same *library calls* the reference makes on in-memory
  arrays (no JPEG I/O); used as the timed CPU baseline and to cross-check the
  spec functions at test time.

Parity pinning: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the oracle is pinned (a) against the reference's own
functions imported from /root/reference in the build container, with outputs
committed under tests/golden/ by tests/golden/make_golden.py, and (b) live
against Pillow/OpenCV/NumPy, which both the build container and the GPU box
carry.
"""
