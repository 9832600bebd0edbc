"""The built library's SASS (no GPU needed): the kernels are hand-written integer / byte code for
sm_100a — no tensor-core instructions, no library kernels — and the match search stages its tile
with a bulk asynchronous copy on an mbarrier."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "raisin_b200", "libraisin_b200.so")


@pytest.fixture(scope="module")
def sass():
    if not shutil.which("cuobjdump") or not os.path.exists(SO):
        pytest.skip("cuobjdump or the built library is missing")
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            funcs[cur].append(ln)
    return out, funcs


def test_built_for_sm100a_only(sass):
    out, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs


def test_every_kernel_is_ours_and_integer(sass):
    _, funcs = sass
    assert len(funcs) > 60
    for name, body in funcs.items():
        assert "rsn" in name, name  # no library kernels linked in
        text = "\n".join(body)
        assert not re.search(r"\b(HMMA|IMMA|UTCMMA|UTCHMMA)\b", text), name  # nothing here is a dense contraction


def test_match_search_stages_its_tile_with_a_bulk_copy(sass):
    _, funcs = sass
    for key in ("12k_match_tileEPKhmjPjm", "13kb_match_tileEPKNS_6LzFileEPjm"):
        body = "\n".join(next(v for k, v in funcs.items() if key in k))
        assert "UBLKCP" in body and "SYNCS" in body, key
