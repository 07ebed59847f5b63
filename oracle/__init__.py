"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the timed CPU baseline.
The product path (``lct-vqa_b200/``) never imports this package and fails loudly
when its CUDA library is missing.
"""
