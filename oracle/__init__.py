"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  PARITY UNPINNED against the upstream binary:
see the header of ``oracle/oracle.cpp`` and DESIGN.md.  ``oracle/bethe.py`` holds an answer that is
independent of the restatement as well: the exact Bethe-ansatz ground-state energy of the Heisenberg ring.
"""
