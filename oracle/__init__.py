"""CPU oracle for the vsc2022 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  The product
package ``vsc2022_b200`` never imports it and fails loudly when its CUDA
library is missing.

Contents
--------
``shims/``            stand-ins for the three third-party engines the reference
                      imports but that cannot be installed here (faiss, vcsl,
                      matplotlib); with ``shims/`` first on ``sys.path`` the
                      UNMODIFIED reference package under ``/root/reference``
                      imports and all 20 of its unit tests run
                      (``oracle/run_reference_tests.py``).
``tn_networkx.py``    literal restatement of VCSL ``tn`` on networkx (authority
                      for stage C).
``tn_fast.c``         array formulation of the same algorithm in plain C
                      (validated against ``tn_networkx`` on thousands of seeded
                      cases; used for full-size parity checks).
``search_numpy.py``   numpy restatement of vsc.index / vsc.candidates /
                      score_normalization (stage B), self-contained so it can
                      travel to the GPU box where ``/root/reference`` is absent.
``localize_numpy.py`` restatement of vsc.baseline.localization (stage C glue).
``make_golden.py``    runs the unmodified reference over the shims on seeded
                      inputs and freezes inputs+outputs under ``tests/golden``.

Parity status: see the table in DESIGN.md ("Oracle pinning").
"""
