"""Test infrastructure only: CPU oracle for the pyDEM hot path.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the *checker* (never as the thing measured on the
GPU arm or shipped).  The product (``pydem_b200``) never imports this package.
"""
