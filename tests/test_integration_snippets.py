"""the C++ snippets a MegaPath-Nano maintainer would paste must compile against include/mpn_ssw_batch.h as
written (argument order, types, names) -- the document cannot drift from the ABI.  Syntax check only (g++ -fsyntax-only): no GPU."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = r"""
#include <cstdint>
#include <vector>
#include "mpn_ssw_batch.h"
void snippets(const int8_t* score_matrix, const int8_t* reads, const int64_t* read_off, const int8_t* refs, const int64_t* ref_off, const int32_t* masklen,
              int64_t npairs, int64_t total_read_bases, int64_t total_ref_bases, const int8_t* arena, int64_t arena_bytes, const int64_t* rd_start,
              const int32_t* rd_len, const int64_t* rf_start, const int32_t* rf_len, uint8_t* reads4, uint8_t* refs4, uint8_t* reads2, uint8_t* refs2)
{
"""


def test_batched_submit_snippets_compile(tmp_path):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 3b."):text.index("## 4.")]
    blocks = re.findall(r"```cpp\n(.*?)```", sec, flags=re.S)
    assert len(blocks) >= 4, "section 3b lost its snippets"
    body = "\n".join(b.replace('#include "mpn_ssw_batch.h"', "") for b in blocks)
    src = tmp_path / "snippets.cpp"
    src.write_text(PRELUDE + body + "\n(void)pool;\n}\n")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:3000]
