"""Every reference citation (file:line) in the boundary header, the integration notes and the design doc
must point at a real line of the reference tree.  Runs only where /root/reference is mounted (the build
container); skipped on the GPU box."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
DOCS = ["include/pbsgpu.h", "include/pbsgpu.hpp", "INTEGRATION.md", "DESIGN.md", "go/gpuchunk/gpuchunk.go",
        "pbs_plus_b200/transfer.py", "pbs_plus_b200/buzhash.py", "oracle/oracle.c"]
ALIASES = {"commit.go": "internal/pxarmount/commit.go", "log_cleanup.go": "internal/server/backup/log_cleanup.go",
           "format.go": "internal/pxar/format.go", "command.go": "internal/server/backup/command.go"}
CITE = re.compile(r"((?:[\w./-]+/)?[\w.-]+\.(?:go|md|yaml|mod)):(\d+)(?:-(\d+))?((?:,\s?:?\d+(?:-\d+)?)*)")


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted here")
def test_reference_citations_resolve():
    checked = 0
    for doc in DOCS:
        text = (ROOT / doc).read_text()
        for m in CITE.finditer(text):
            name = m.group(1)
            if name.startswith("tests/") or name.startswith("profiles/") or name.startswith("pbs_plus_b200/"):
                continue
            path = REF / name if (REF / name).exists() else REF / ALIASES.get(name.split("/")[-1], name)
            if not path.exists():
                cands = list(REF.rglob(name.split("/")[-1]))
                assert len(cands) >= 1, f"{doc}: cited file {name} not in the reference"
                path = cands[0]
            n_lines = sum(1 for _ in path.open(errors="replace"))
            nums = [int(m.group(2))] + ([int(m.group(3))] if m.group(3) else [])
            nums += [int(x) for x in re.findall(r"\d+", m.group(4) or "")]
            for ln in nums:
                assert 1 <= ln <= n_lines, f"{doc}: {name}:{ln} is beyond the file's {n_lines} lines"
            checked += 1
    assert checked >= 25
