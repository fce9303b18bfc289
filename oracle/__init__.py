"""oracle — CPU checkers for the RPEFlow cost-volume hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package.  Nothing under ``rpeflow_b200/`` does (tests/test_boundary.py greps for it).

Two checkers live here:

* ``oracle.spec``  — numpy front-end of ``oracle/spec.c`` (scalar C restatement, exact stated rules for the
  index ops).  Parity: pinned against ``tests/golden/*.npz`` (fixtures produced from the unmodified
  reference by ``tests/golden/make_golden.py``).
* ``oracle.torch_ref`` — the same path re-stated with torch CPU ops, i.e. the arithmetic the reference's
  own fallbacks execute (``models/csrc/wrapper.py``, ``models/pwc3d_core.py``, ``models/utils.py``,
  ``event_utils.py``, ``dsec.py``).  This is what ``bench.py --impl reference`` times.

``oracle.refcuda`` (optional) opens ``oracle/_ref/libref_kernels.so`` — the reference's own CUDA kernels
compiled from /root/reference for sm_100a by ``oracle/Makefile`` — as a second, GPU-side checker.
"""
