"""The run-time compiled order-0 kernel without a device: phyfoo_jit_check generates the straight-line pass of a tree and
compiles it with NVRTC for sm_100a (no GPU needed for that); csrc/jit_embed.inc must be the current text of the headers the
generated source includes."""
import ctypes as C
import os

import numpy as np
import pytest

from util import ROOT

from phyfoo_b200 import _abi, lib as plib, synth
from phyfoo_b200.newick import parse_newick


def _check(newick, W, gaps):
    tree = parse_newick(newick)
    pwm = synth.random_pwm(W, 0, 1)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, np.tile(tree.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    src = C.create_string_buffer(1 << 20)
    log = C.create_string_buffer(1 << 16)
    n = C.c_int64()
    rc = plib.lib().phyfoo_jit_check(C.byref(desc), int(gaps), src, len(src), log, len(log), C.byref(n))
    return rc, src.value.decode(), log.value.decode(), n.value, tree


def test_embedded_headers_are_current():
    assert open(plib.JIT_EMBED).read() == plib.jit_embed_text(), "run phyfoo_b200.lib.build(): csrc/jit_embed.inc is stale"


@pytest.mark.parametrize("newick,W", [(synth.star_newick(5), 7), (synth.caterpillar_newick(5), 10),
                                      (synth.random_binary_newick(12, 0x5EED0003), 12), (synth.random_binary_newick(30, 0x5EED0005), 12),
                                      ("((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)", 33)])
@pytest.mark.parametrize("gaps", [False, True])
def test_generated_kernel_compiles_for_sm100a(newick, W, gaps):
    rc, src, log, cubin_bytes, tree = _check(newick, W, gaps)
    if rc == _abi.EUNSUPPORTED and "nvrtc" in log.lower():
        pytest.skip("no NVRTC library in this environment: " + log)
    assert rc == 0, log[:2000]
    assert cubin_bytes > 10000
    n_internal = tree.n_nodes - tree.n_species
    assert src.count("// internal node ") == n_internal                       # one straight-line block per hidden node
    assert f"#define JW {W}\n" in src and 'extern "C" __global__' in src
    assert ("jit_gap_leaf" in src) == gaps
    # every leaf is read exactly once per pass: one literal shift per species
    for sp in range(tree.n_species):
        assert src.count(f"(cw.b >> {2 * sp}) & 3u") == 1
