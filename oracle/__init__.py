"""ORACLE - test infrastructure only.

CPU restatement of the reference hot path (``port.py``: PyTorch-CPU ops in the reference's order;
``csr_ref.c``: plain C for the integer / bit-exact collate + CSR + degree part).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it - as the checker or the timed CPU baseline, never as a product path.
"""
