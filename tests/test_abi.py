"""C-ABI checks that need no GPU: the library loads, exports every symbol
include/sylver_b200.h declares, and its public structs are laid out exactly as
the reference's (/root/reference/include/sylver/sylver.h:19-71,
src/sylver_ciface.hxx:39-87)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import sylver_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sylver_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:sylver|spldlt)_[A-Za-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(lib):
    names = _declared_functions()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/sylver_b200.h but not exported"
    for nm in sb.EXPORTS:
        assert nm in names, f"{nm} bound by the Python layer but not declared in the header"


def test_struct_layouts_match_header(lib):
    """Compile a tiny C program against the header and compare sizeof/offsetof with
    the ctypes mirrors (which follow the reference struct definitions field by field)."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "sylver_b200.h"
int main(void){
 printf("%zu %zu %zu %zu\n", sizeof(sylver_inform_t), sizeof(sylver_options_t), sizeof(sylver_options_c), sizeof(sylver_inform_c));
 printf("%zu %zu %zu %zu %zu\n", offsetof(sylver_inform_t,num_factor), offsetof(sylver_inform_t,num_neg), offsetof(sylver_inform_t,unused),
        offsetof(sylver_options_t,min_gpu_work), offsetof(sylver_options_t,gpu_perf_coeff));
 printf("%zu %zu %zu\n", offsetof(sylver_options_c,small), offsetof(sylver_options_c,nb), offsetof(sylver_inform_c,not_second_pass));
 return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    vals = list(map(int, out))
    assert vals[0] == C.sizeof(sb.Inform)
    assert vals[1] == C.sizeof(sb.Options)
    assert vals[2] == C.sizeof(sb.OptionsC)
    assert vals[3] == C.sizeof(sb.InformC)
    assert vals[4] == sb.Inform.num_factor.offset
    assert vals[5] == sb.Inform.num_neg.offset
    assert vals[6] == sb.Inform.unused.offset
    assert vals[7] == sb.Options.min_gpu_work.offset
    assert vals[8] == sb.Options.gpu_perf_coeff.offset
    assert vals[9] == sb.OptionsC.small.offset
    assert vals[10] == sb.OptionsC.nb.offset
    assert vals[11] == sb.InformC.not_second_pass.offset


def test_default_options_are_the_reference_defaults(lib):
    # /root/reference/src/sylver_datatypes_mod.F90:97-198
    o = sb.Options()
    lib.sylver_default_options(C.byref(o))
    assert (o.ordering, o.nemin, o.prune_tree, o.min_gpu_work) == (1, 32, True, 5 * 10 ** 9)
    assert (o.scaling, o.pivot_method, o.small, o.u) == (0, 2, 1e-20, 0.01)
    assert (o.small_subtree_threshold, o.nb, o.action, o.use_gpu) == (4 * 10 ** 6, 256, True, True)
    assert (o.gpu_perf_coeff, o.failed_pivot_method) == (1.0, 1)


def test_version_and_device_count_do_not_need_a_gpu(lib):
    assert b"sm_100a" in lib.sylver_b200_version()
    assert lib.sylver_b200_device_count() >= 0


def test_product_does_not_link_the_oracle():
    out = subprocess.run(["ldd", sb.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in out and "openblas" not in out


def _build_example(tmpdir):
    exe = os.path.join(tmpdir, "example")
    libdir = os.path.join(ROOT, "sylver_b200")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "spldlt_simple_example.c"), "-L", libdir, "-lsylver_b200",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def test_c_example_compiles_and_links(lib):
    """The reference's C example, with a supplied order, builds against the header and the
    library with a plain C compiler (no C++, no CUDA headers needed by the caller)."""
    with tempfile.TemporaryDirectory() as d:
        exe = _build_example(d)
        assert os.path.exists(exe)
        if sb.device_count() == 0:
            # without a device the program must fail loudly at factorize, not fall back
            r = subprocess.run([exe], capture_output=True, text=True)
            assert r.returncode == 1 and "factorize failed: -51" in r.stdout
