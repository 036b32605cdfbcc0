"""TEST INFRASTRUCTURE: CPU restatement of the reference's algorithm for the retrieval hot path.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The
product package (lightningdot_b200/) must never import anything from here.
"""
