"""TEST INFRASTRUCTURE ONLY - CPU oracle for the NMF multiplicative-update path.

Nothing in the product package (``pymf_b200``) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs are allowed to.
"""
