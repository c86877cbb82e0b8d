"""The C-ABI libraries load and export every symbol the headers declare (no compute calls: no GPU needed)."""
import ctypes
import os
import re

import support as S

b2pt = S.b2pt


def declared(header, prefix):
    text = open(os.path.join(S.ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, text)))


def test_gpu_library_exports():
    names = declared("b2pt.h", "b2pt_")
    assert len(names) >= 25
    lib = ctypes.CDLL(b2pt.GPU_LIB)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.b2pt_abi_version.restype = ctypes.c_int
    assert lib.b2pt_abi_version() == 1


def test_host_library_exports():
    names = [n for n in declared("b2pt_host.h", "b2pt_host_")]
    assert len(names) >= 35
    lib = ctypes.CDLL(b2pt.HOST_LIB)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match_header():
    """ctypes mirrors of the POD structs have the C sizes (compiled probe)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "b2pt.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",sizeof(b2pt_material),sizeof(b2pt_node),' \
          'sizeof(b2pt_scene_desc),sizeof(b2pt_camera),sizeof(b2pt_render_params),sizeof(b2pt_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "p.c")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(S.ROOT, "include"), c, "-o", os.path.join(td, "p")], check=True)
        out = subprocess.run([os.path.join(td, "p")], capture_output=True, text=True, check=True).stdout.split()
    want = [ctypes.sizeof(t) for t in (b2pt.Material, b2pt.Node, b2pt.SceneDesc, b2pt.Camera, b2pt.RenderParams, b2pt.Stats)]
    assert [int(x) for x in out] == want


def test_no_cpu_fallback_without_gpu():
    """Creating a context on a machine without a usable B200 fails loudly (B2PT_ERR_NO_DEVICE), it never falls back."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        b2pt.Context(0)
    except RuntimeError as e:
        assert "no CUDA device" in str(e) or "CPU path" in str(e) or "failed" in str(e)
    else:
        raise AssertionError("Context(0) succeeded without a GPU")
