import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    from tests.util import Golden
    return Golden()


def pytest_sessionfinish(session, exitstatus):
    """Parity report of a GPU session: per comparison the fraction of elements outside the PLAIN north-star bound."""
    from tests import util
    if not util.PARITY_REPORT:
        return
    import json
    out = os.path.join(REPO, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as fh:
            json.dump({"max_frac_outside_plain_bound": max(r["frac_outside_plain_bound"] for r in util.PARITY_REPORT),
                       "max_mixed_err": max(r["max_mixed_err"] for r in util.PARITY_REPORT),
                       "comparisons": util.PARITY_REPORT}, fh, indent=1)
    except OSError:
        pass
