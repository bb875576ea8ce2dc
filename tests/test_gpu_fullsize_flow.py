"""The full-size parity procedure (tools/fullsize.py) on a reduced genome: index built by dartgpu_index_build, the
canonical reference multithreaded (records in completion order), dart_b200_map in input order; the SAM records must be
the same multiset (tools/samhash.cpp) and junctions.tab identical. The recorded full-size runs (3.1 Gbp, 20 M pairs) are
in profiles/r01_fullsize_config2_config3.json; this keeps the procedure itself under test."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, need_ref

pytestmark = pytest.mark.gpu


def test_fullsize_procedure_on_a_reduced_genome(tmp_path):
    need_ref()
    out = tmp_path / "summary.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fullsize.py"), "--scale", "0.004", "--pairs2", "60000",
                    "--pairs3", "6000", "--dir", str(tmp_path / "work"), "--out", str(out)], check=True,
                   stdout=subprocess.DEVNULL)
    s = json.load(open(out))
    for cfg in ("config2", "config3"):
        assert s[cfg]["sam_identical_multiset"] and s[cfg]["junctions_identical"], s[cfg]
        assert s[cfg]["gpu_hash"][0] == 2 * s[cfg]["pairs"]
    assert s["config2"]["junction_lines"] > 0
