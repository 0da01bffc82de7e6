"""The C-ABI library loads and exports every symbol include/svit_b200.h declares (no compute calls)."""
import os
import re

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "svit_b200.h")).read()
    return sorted(set(re.findall(r"^int\s+(svit_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_the_hot_path():
    names = _declared()
    for must in ("svit_pool_ln_fwd", "svit_pool_ln_bwd", "svit_attn_fwd", "svit_attn_bwd", "svit_gemm",
                 "svit_layernorm_fwd", "svit_skip_maxpool_fwd", "svit_assemble_tokens_fwd", "svit_roi_tokens_fwd",
                 "svit_match_haog", "svit_im2col3d", "svit_gather_cls_obj_fwd"):
        assert must in names


def test_library_builds_loads_and_exports_everything():
    from svit_b200 import _lib, build

    build.build()
    L = _lib.lib()
    declared = _declared()
    assert sorted(_lib.PROTOTYPES) == declared
    for name in declared:
        assert hasattr(L, name), name
    assert L.svit_abi_version() == 100


def test_struct_layouts_match_header():
    # field order of the ctypes mirrors must follow the header's structs
    from svit_b200 import _lib

    src = open(os.path.join(ROOT, "include", "svit_b200.h")).read()
    for cname, cls in (("svit_gemm_args", _lib.GemmArgs), ("svit_attn_args", _lib.AttnArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1].lstrip("*")
            fields.append(first)
            fields += [n.strip().lstrip("*") for n in names[1:]]
        assert fields == [f[0] for f in cls._fields_], cname


def test_sass_is_sm100a():
    import subprocess

    from svit_b200 import _lib

    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
