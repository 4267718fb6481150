"""The C-ABI library loads, exports every symbol include/phyfoo.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from util import ROOT

from phyfoo_b200 import _abi
from phyfoo_b200 import lib as plib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "phyfoo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phyfoo_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol():
    plib.build()
    L = plib.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"libphyfoo.so does not export {s}"
    assert sorted(plib.SYMBOLS) == syms


def test_integration_table_names_every_symbol():
    """INTEGRATION.md's appendix says, entry point by entry point, which reference code each one stands in for."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    appendix = text[text.index("## Appendix"):]
    missing = [s for s in header_symbols() if not re.search(r"`" + s + r"`", appendix)]
    assert not missing, f"INTEGRATION.md appendix does not name {missing}"


def _java_type(c_decl):
    c = c_decl.strip()
    if "*" in c:
        return "ADDRESS"
    for prefix, j in (("double", "JAVA_DOUBLE"), ("int64_t", "JAVA_LONG"), ("float", "JAVA_FLOAT"), ("int32_t", "JAVA_INT"), ("int", "JAVA_INT")):
        if c.startswith(prefix):
            return j
    raise AssertionError(f"no Java layout for {c_decl!r}")


def test_panama_stubs_match_header():
    """Every downcall handle INTEGRATION.md shows has the argument and return layouts of its prototype in include/phyfoo.h."""
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "phyfoo.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b([a-z_0-9\* ]+?)\b(phyfoo_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(3).split(",") if a.strip() and a.strip() != "void"]
        protos[m.group(2)] = (m.group(1).strip(), args)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stubs = list(re.finditer(r'h\("(phyfoo_[a-z0-9_]+)",\s*FunctionDescriptor\.(ofVoid|of)\(([^;]*?)\)\)\s*;', doc, flags=re.S))
    assert len(stubs) >= 40
    for m in stubs:
        name, kind = m.group(1), m.group(2)
        got = [a.strip() for a in re.sub(r"\s+", " ", m.group(3)).split(",") if a.strip()]
        ret, args = protos[name]
        want = [_java_type(a) for a in args]
        if kind == "of":
            want = [_java_type(ret + " ")] + want
        else:
            assert ret == "void", name
        assert got == want, f"{name}: stub {got} != header {want}"


def test_java_model_desc_layout_follows_the_header_struct():
    """INTEGRATION.md's MODEL_DESC StructLayout: the header's fields in order with their Java layouts, and padding exactly
    where natural C alignment puts it (an int before an 8-byte member at an offset that is not a multiple of 8)."""
    text = open(os.path.join(ROOT, "include", "phyfoo.h")).read()
    body = text[text.index("typedef struct {"):text.index("} phyfoo_model_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [(m.group(1).strip(), m.group(2)) for m in re.finditer(r"\n\s*([a-z_0-9 ]+\*?)\s*\b([a-z_]+);", body)]
    assert [f for _, f in fields][:4] == ["kind", "width", "order", "n_nodes"] and len(fields) == 15
    want, off = [], 0
    for decl, name in fields:
        j = _java_type(decl + " ")
        size = 4 if j == "JAVA_INT" else 8
        if off % size:
            want.append(("pad", size - off % size))
            off += size - off % size
        want.append((j, name))
        off += size
    assert off == C.sizeof(_abi.ModelDesc)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("static final StructLayout MODEL_DESC"):]
    block = block[:block.index(";")]
    got = []
    for m in re.finditer(r'(JAVA_INT|JAVA_DOUBLE|JAVA_LONG|ADDRESS)\.withName\("([a-z_]+)"\)|MemoryLayout\.paddingLayout\((\d+)\)', block):
        got.append(("pad", int(m.group(3))) if m.group(3) else (m.group(1), m.group(2)))
    assert got == want


def test_struct_layouts_match_header():
    # phyfoo_model_desc: 4 int32, 2 pointers, int32 (+pad), double, 2 pointers, 4 int32, double
    assert C.sizeof(_abi.ModelDesc) == 16 + 16 + 8 + 8 + 16 + 16 + 8
    assert C.sizeof(_abi.Stats) == 4 * 8 + 4 + 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = plib.lib().phyfoo_init(0, C.byref(h))
    assert rc == _abi.ENODEVICE
    assert b"no CPU fallback" in plib.lib().phyfoo_last_error(None)
    from phyfoo_b200 import core
    with pytest.raises(plib.PhyfooError):
        core.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "phyfoo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("test infrastructure", "").lower() or f in (), \
                    f"{f} mentions the oracle: the product path must not depend on it"


def test_header_is_plain_c_and_a_c_caller_links(tmp_path):
    """include/phyfoo.h compiles as pedantic C99 (what cgo / JNI / Panama's jextract read), a C caller links against
    libphyfoo.so with nothing else on the link line, and without a device both init calls fail loudly with PHYFOO_ENODEVICE."""
    import subprocess
    plib.build()
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include "phyfoo.h"
#include <stdio.h>
int main(void) {
    phyfoo_ctx* ctx = 0;
    phyfoo_mctx* mctx = 0;
    int dev[1] = {0};
    int rc = phyfoo_init(0, &ctx);
    int rcm = phyfoo_init_multi(dev, 1, &mctx);
    printf("%d %d %d|%s\n", rc, rcm, (int)sizeof(phyfoo_model_desc), rc ? phyfoo_last_error(ctx) : "");
    if (mctx) phyfoo_destroy_multi(mctx);
    if (ctx) phyfoo_destroy(ctx);
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.join(ROOT, "phyfoo_b200")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-l:libphyfoo.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    head, _, err = out.stdout.strip().partition("|")
    rc, rcm, size = (int(x) for x in head.split())
    assert size == C.sizeof(_abi.ModelDesc)
    assert (rc, rcm) in ((_abi.OK, _abi.OK), (_abi.ENODEVICE, _abi.ENODEVICE))
    if rc:
        assert "no CPU fallback" in err
