"""ORACLE / TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.  PARITY UNPINNED.

A CPU restatement, shaped like the ``norse`` package, of the four Norse 0.0.7
primitives the reference's spiking heads call (README.md:13 pins norse==0.0.7;
the package is neither vendored under /root/reference nor installable here).
It exists so that /root/reference/rpn.py and /root/reference/faster_rcnn.py can
be imported UNMODIFIED to generate golden vectors (oracle/gen_golden.py).

"Parity unpinned": the reference ships no tests or golden vectors for this
path and the Norse sources are not available offline, so the equations below
are restated from the published Norse 0.0.7 algorithm
and pinned only by closed-form known answers (tests/test_oracle_known_answers.py)
and by the reference's own call sites.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything under oracle/.
"""
