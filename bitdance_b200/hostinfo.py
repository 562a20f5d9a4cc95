"""Facts about the host this process runs on (no torch, no CUDA)."""
from __future__ import annotations


def usable_cpus(cgroup_root: str = "/sys/fs/cgroup") -> int:
    """Host threads this process may really use: cpu_count, limited by the affinity mask and the cgroup CPU quota (an
    OpenMP team wider than the quota spins on its barriers and runs orders of magnitude slower)."""
    import math
    import os
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in (os.path.join(cgroup_root, "cpu.max"), os.path.join(cgroup_root, "cpu", "cpu.cfs_quota_us")):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):                       # cgroup v2: "<quota|max> <period>"
                if txt[0] != "max":
                    n = min(n, max(1, math.floor(int(txt[0]) / int(txt[1]))))
            else:                                              # cgroup v1
                q = int(txt[0])
                if q > 0:
                    per = int(open(os.path.join(cgroup_root, "cpu", "cpu.cfs_period_us")).read())
                    n = min(n, max(1, q // per))
        except (OSError, ValueError, IndexError, ZeroDivisionError):
            continue
    return max(1, n)
