"""oracle/ — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's algorithms).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker.  Nothing under pcfa_b200/ imports it.

Parity pinning (the reference has no tests or golden vectors): the restatements
are pinned against (1) outputs of the reference code executed in the authoring container
(oracle/make_golden.py → tests/golden/*.npz, committed) and (2) the reference's own spatial
correlation sampler CPU extension compiled from its sources (oracle/build_ref.py → oracle/_ref/).
"""
