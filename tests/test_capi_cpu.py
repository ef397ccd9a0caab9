"""CPU-only checks of the product library: it loads, exports every symbol pbsgpu.h
declares, and its host-side logic (config derivation, default table, error paths)
matches the oracle.  No compute calls (there is no GPU here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
import pbs_plus_b200 as pg
from pbs_plus_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    hdr = (ROOT / "include" / "pbsgpu.h").read_text()
    declared = set(re.findall(r"\b(pbsgpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.pbsgpu_version() == 201


def test_config_matches_oracle_and_rejects_bad_sizes():
    for avg in (256, 1024, 4096, 1 << 20, 4 << 20, 1 << 29):
        a, b = pg.make_config(avg), oracle.config(avg)
        for f in ("avg", "min", "max", "mask", "break_min", "window"):
            assert getattr(a, f) == getattr(b, f)
        assert list(a.table) == list(b.table)
    for bad in (0, 100, 255, 3 << 20, 1 << 30):
        with pytest.raises(pg.PbsGpuError):
            pg.make_config(bad)
    c = pg.buzhash.NewConfig(4096)                      # the reference's call (commit.go:303)
    assert (c.avg, c.min, c.max, c.mask) == (4 << 20, 1 << 20, 16 << 20, 0x7FFFFF)
    with pytest.raises(ValueError):
        pg.buzhash.NewConfig(0)


def test_product_table_copy_equals_oracle_table():
    assert (pg.default_table() == oracle.default_table()).all()
    t = ((np.arange(256, dtype=np.uint64) * 2654435761) % (2**32)).astype(np.uint32)
    assert list(pg.make_config(1024, t).table) == t.tolist()


def test_struct_layouts_match_between_oracle_and_product():
    assert C.sizeof(_lib.Cfg) == C.sizeof(oracle.Cfg) == 6 * 4 + 1024
    assert C.sizeof(_lib.Chunk) == C.sizeof(oracle.Chunk) == 48
    assert C.sizeof(_lib.Corpus) == C.sizeof(oracle.Corpus)
    assert _lib.CHUNK_DTYPE == oracle.CHUNK_DTYPE


def test_open_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pg.PbsGpuError) as e:
        pg.Engine()
    assert e.value.code == _lib.ENODEV


def test_product_never_imports_the_oracle():
    for p in (ROOT / "pbs_plus_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".inc", ".cpp", ".h", ".hpp"):
            txt = p.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt, p
            assert '#include "../../oracle' not in txt and "liboracle" not in txt, p


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: pbsgpu.h must compile as C11 (what cgo feeds to its C compiler) and a C
    program must link against the library and call the host-only entry points."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include "pbsgpu.h"
int main(void) {
    pbsgpu_cfg c;
    if (pbsgpu_config_kib(4096, NULL, &c) != 0) return 1;
    if (c.avg != (4u << 20) || c.min != (1u << 20) || c.max != (16u << 20) || c.mask != 0x7FFFFFu) return 2;
    if (pbsgpu_config(3000, NULL, &c) != PBSGPU_EINVAL) return 3;
    if (pbsgpu_default_table()[0] != 0x458be752u) return 4;
    if (pbsgpu_didx_size(10) != 4096 + 400) return 5;
    unsigned char h[12]; pbsgpu_blob_header(0x11223344u, h);
    if (h[0] != 66 || h[8] != 0x44 || h[11] != 0x11) return 6;
    printf("version %d\n", pbsgpu_version());
    return 0;
}
''')
    exe = tmp_path / "t.bin"
    lib = ROOT / "pbs_plus_b200"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-o", str(exe), str(src),
                           f"-L{lib}", "-lpbsgpu", f"-Wl,-rpath,{lib}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "version 201", (out.returncode, out.stdout, out.stderr)
