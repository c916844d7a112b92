"""Shared fixtures: re-export of the package's synthetic workloads."""

from hy_b200.workloads import *  # noqa: F401,F403
from hy_b200.workloads import OSS_MASSES, OSS_G, OSS_IC, PEND_IC, CR3BP_IC  # noqa: F401
