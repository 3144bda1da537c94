"""CPU oracle for the DDPM denoising hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the
CPU baseline being timed.  See ``oracle/msd_oracle.py`` for the parity status.
"""
