"""The C ABI without Python in the call path: tests/c_abi/abi_check.c is compiled with gcc, dlopens libmcmcdate_b200.so and runs
mcd_create -> mcd_eval_grad / mcd_eval -> mcd_destroy on the reference's 12-leaf data set stored as plain numbers
(tests/golden/abi_case_12_leaves.txt), comparing with the oracle's values in the same file."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_check.c")
LIB = os.path.join(ROOT, "mcmc-date_b200", "libmcmcdate_b200.so")
CASE = os.path.join(ROOT, "tests", "golden", "abi_case_12_leaves.txt")


def _build(tmp_path):
    exe = str(tmp_path / "abi_check")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-o", exe, SRC, "-ldl", "-lm"])
    return exe


def test_c_program_resolves_the_entry_points(tmp_path):
    """no GPU: the header compiles as C, the library loads and exports what the program binds"""
    r = subprocess.run([_build(tmp_path), LIB, "--symbols"], capture_output=True, text=True)
    assert r.returncode == 0 and "symbols ok" in r.stdout, r.stderr


@pytest.mark.gpu
def test_c_program_evaluates_the_fixture(tmp_path):
    r = subprocess.run([_build(tmp_path), LIB, CASE], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout, r.stderr)


def test_documented_struct_layout(tmp_path):
    """INTEGRATION.md's Haskell marshalling pokes mcd_model_desc by byte offset: check every offset it uses (and the size)
    against the header with a C program"""
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("withModelDesc ::"):text.index('foreign import ccall safe "mcd_create"')]
    used = sorted({int(m) for m in re.findall(r"pokeByteOff d\s+(\d+)", block)})
    assert "allocaBytesAligned 216 8" in block and len(used) >= 22
    fields = ["n_nodes", "parent", "clock_model", "likelihood", "mean", "precision", "logdet_sigma", "ht", "n_cal", "cal_node",
              "cal_lo", "cal_lo_p", "cal_hi", "cal_hi_p", "n_con", "con_young", "con_old", "con_p", "n_brace", "brace_off",
              "brace_node", "brace_sd", "device", "max_batch", "precision_chol", "n_sparse", "sparse_row", "sparse_col", "sparse_val"]
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){%s printf("%%zu\\n", sizeof(mcd_model_desc)); return 0;}\n'
                   % (os.path.join(ROOT, "include", "mcmcdate_b200.h"),
                      "".join('printf("%%zu\\n", offsetof(mcd_model_desc, %s));' % f for f in fields)))
    exe = str(tmp_path / "off")
    subprocess.check_call(["gcc", "-o", exe, str(src)])
    vals = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    offsets, size = vals[:-1], vals[-1]
    assert size == 216
    documented_fields = fields[:24]                      # everything up to max_batch is poked; the rest stays zero
    assert used == sorted(offsets[:24]), (used, offsets[:24])
    assert offsets[24] == 176 and offsets[25] == 184     # mentioned in the comment of the shim
