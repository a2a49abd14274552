"""Guards against a documentation-generation slip (round 1 turned BASELINE.md into 3.3 MB of repeated table rows):
hand-written docs stay small and are not dominated by one repeated line."""
import collections
import os

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["BASELINE.md", "DESIGN.md", "INTEGRATION.md", "README.md", "profiles/README.md"]


@pytest.mark.parametrize("name", DOCS)
def test_doc_is_sane(name):
    path = os.path.join(REPO, name)
    assert os.path.exists(path), name
    text = open(path, encoding="utf-8").read()
    assert len(text) < 200_000, f"{name} is {len(text)} bytes"
    lines = [ln for ln in text.splitlines() if len(ln.strip()) > 20]
    if lines:
        top, n = collections.Counter(lines).most_common(1)[0]
        assert n <= max(4, len(lines) // 20), f"{name}: the line {top[:60]!r} repeats {n} times"


def test_baseline_md_keeps_its_sections():
    text = open(os.path.join(REPO, "BASELINE.md"), encoding="utf-8").read()
    for head in ("## 1. Published reference numbers", "## 2. Hardware ceilings", "## 3. CPU-baseline plan", "## 4. Measured"):
        assert head in text
