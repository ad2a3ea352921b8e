"""CPU oracles for the KGDet point-set head hot path.

TEST INFRASTRUCTURE ONLY: importable from ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product
package ``kgdet_b200`` never imports this package (tests/test_boundary_cpu.py
greps for it).
"""
