"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the gypsum acquisition / tracking hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``gypsum_b200``) never does.
"""
