"""The plain-C oracle ports against committed golden digests of the REFERENCE's outputs (tests/golden/digests.json, written by
tests/golden/make_digests.py from oracle/_ref in the build container).  Needs neither /root/reference nor oracle/_ref, so the pin
travels with the repository."""
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_digests  # noqa: E402

DIGESTS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "digests.json")))


@pytest.mark.parametrize("name", sorted(DIGESTS))
def test_port_reproduces_reference_digest(name):
    fn = make_digests.cases("port")[name]
    assert make_digests.digest(fn()) == DIGESTS[name]
