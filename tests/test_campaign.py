"""Two seeds of tools/fuzz_campaign.py per test run (reference binary vs Python oracle vs the product's host layer, with random
option vectors, the generator's edge mode and the getsv -F / -B variants): keeps the campaign runnable. The seeds are fixed so that
the suite is deterministic; SEEKSV_B200_FRESH_SEEDS=1 takes two seeds that change with the day instead. CPU only; needs
oracle/_ref (built by oracle/build_ref.sh where /root/reference exists; travels to the GPU box)."""
import os
import subprocess
import sys
import time

import pytest

from conftest import ROOT

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "seeksv")), reason="no oracle/_ref")


@pytest.mark.parametrize("extra", [["--options", "1"], ["--edge", "--connect", "60"]])
def test_two_fresh_seeds(extra):
    first = 100000 + 2 * (int(time.time()) // 86400 % 10000) if os.environ.get("SEEKSV_B200_FRESH_SEEDS") else 7000
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_campaign.py"), "--seeds", "%d:%d" % (first, first + 2),
                        "--records", "1200"] + extra, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 of 2 seeds differed" in r.stdout
