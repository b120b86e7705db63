"""lfsd_b200 — B200-native CPDP (Learning from Sparse Demonstrations) gradient-iteration path.

Host API mirrors the reference (``/root/reference/CPDP/CPDP.py``, ``/root/reference/JinEnv/JinEnv.py``):
``CPDP.COCSys`` with ``setDyn/setPathCost/setFinalCost/cocSolver/diffPMP/auxSysSolver`` plus batched variants,
``JinEnv`` model definitions, and the CasADi-like symbolic names in ``sx``.  All numerics run in hand-written
sm_100a CUDA kernels reached through a C-ABI shared library (``include/cpdp.h``); there is no CPU fallback.
"""
from . import sx          # noqa: F401
from . import JinEnv      # noqa: F401
