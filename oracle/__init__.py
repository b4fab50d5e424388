"""CPU ORACLE for the contact-map hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  The product
package ``contact_map_b200`` never imports it and has no CPU fallback.

Contents
--------
``cm_oracle.c`` / ``clib.py``  restated MDTraj float32 arithmetic (neighbour list,
                               distances) + closed-form contact enumeration.
``refpath.py``                 plain-Python restatement of the reference's
                               per-frame set/Counter algorithm (the CPU baseline).
``mdtraj/``                    stand-in for the absent third-party MDTraj so the
                               UNMODIFIED reference can run in the build container
                               (golden-vector generation, oracle validation).

Parity status: combinatorics pinned by the reference's golden vectors and by
outputs of the reference itself (tests/golden); the ulp-level arithmetic is a
restatement of MDTraj's published algorithm ("parity unpinned" at 1 ulp).
"""
