"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the EMLight reference hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or the timed CPU baseline.
The product (``emlight_b200``) never imports this package and raises if its CUDA
library is missing.

Pinning: the reference ships no tests, golden vectors or fixtures (SURVEY.md
section 8c).  Every restatement here is therefore pinned against the reference's own
Python modules executed in the build container (``oracle/make_golden.py`` imports
them from /root/reference and writes ``tests/golden/*.npz``); the CPU test-suite
re-checks each restatement against those committed vectors.
"""
