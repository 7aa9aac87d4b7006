"""TEST INFRASTRUCTURE: CPU oracle for the rSVD hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this package; the product package must not."""
