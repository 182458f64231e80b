"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU (PyTorch-functional) restatement of the reference's algorithm for the training hot path, used as the
parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing in the product package (cbim-medical-image-segmentation_b200/) imports from here.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against the reference ITSELF, imported from /root/reference in the build container by
oracle/make_golden.py, which asserts restatement == reference and writes tests/golden/*.pt.
"""
