import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "reference: needs the unmodified reference (/root/reference, or the copy "
                                               "oracle/make_ref.py ships under oracle/_ref)")


def pytest_collection_modifyitems(config, items):
    import torch

    try:   # GPU boxes show 128 CPUs but grant a 16-CPU cgroup quota (profiles/r02_host_probe.txt): a wider OpenMP team
        from bitdance_b200.hostinfo import usable_cpus   # makes every CPU oracle call crawl
        torch.set_num_threads(usable_cpus())
    except Exception:
        pass

    has_gpu = torch.cuda.is_available()
    has_ref = (os.path.isdir("/root/reference/modeling")
               or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "reference", "modeling")))
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="reference not present (run oracle/make_ref.py in the dev container)")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)
        if "reference" in item.keywords and not has_ref:
            item.add_marker(skip_ref)
