"""CPU oracle for the UMGAP per-read classification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import, link or execute it, and there only as
the checker (or as the timed CPU baseline), never as a fallback for the CUDA
path.

Two restatements of the reference algorithm live here:

* ``oracle/naive`` (this package's Python modules): a line-by-line restatement
  of the reference's Rust (every function cites the reference file:line it
  follows).  Slow, used for small cases, known-answer vectors and as the
  arbiter for the C port.  Where the reference is non-deterministic (HashMap /
  HashSet iteration order in hybrid and MRTL) it returns the
  SET of admissible answers.
* ``oracle/c/umgap_ref.c``: the same algorithms in plain C with pthreads (an
  in-memory FST image walked one node per key byte like ``fst::Map::get``),
  used for larger parity cases and as the timed CPU baseline ("port").

Parity status: the reference is Rust and neither ``cargo`` nor the ``fst``
crate exist in this image, so the reference binary cannot be run here.  The
oracle is pinned against every in-scope unit-test vector and doc-comment example
the reference holds (tests/test_oracle_golden.py).  The one
boundary with no reference-held vector beyond a 2-key round trip is the on-disk
``fst`` 0.3.5 byte format (third-party crate, not vendored): **parity unpinned**
for byte-compatibility with real ``fst``-written files (key/value semantics of
``get`` are pinned by our own writer/reader round trips only).
"""
