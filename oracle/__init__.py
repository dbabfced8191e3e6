"""CPU oracle for the grid->region aggregation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``climate_toolbox_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.
See ``oracle/oracle.py`` for the parity-pinning statement.
"""
from .oracle import *  # noqa: F401,F403
