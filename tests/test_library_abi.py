"""The C-ABI library loads on a CPU-only box and exports every entry point include/puckswarm_b200.h declares.
No compute call is made here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from puckswarm_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "puckswarm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__ as g
        g.build_engine()
    L = C.CDLL(engine.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert set(names) == set(engine.EXPORTS)


def test_struct_layouts_match_the_compiled_header():
    L = engine.load_library()   # load_library itself refuses a layout mismatch
    structs = [abi.ps_config, abi.ps_robot_state, abi.ps_arena_state, abi.ps_localmap, abi.ps_stats, abi.ps_state_view, abi.ps_obs_view]
    for i, t in enumerate(structs):
        assert L.ps_struct_size(i) == C.sizeof(t), t.__name__
    assert L.ps_struct_size(99) == -1


def test_defaults_match_library():
    L = engine.load_library()
    c = abi.ps_config()
    L.ps_config_defaults(C.byref(c))
    d = abi.config_defaults()
    for name, _ in abi.ps_config._fields_:
        assert getattr(c, name) == getattr(d, name), name


def test_no_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine(abi.config_defaults())
    assert ei.value.code in (abi.PS_E_NODEVICE, abi.PS_E_CUDA)


def test_java_config_image_matches_ps_config():
    """java/puckswarm/b200/ArenaB200.java: Config.toBuffer() writes ps_config field by field; the sequence of putInt / putFloat
    must be the header's sequence of int32 / float fields (no JDK here: checked on the source)."""
    src = open(os.path.join(ROOT, "java", "puckswarm", "b200", "ArenaB200.java")).read()
    body = src[src.index("public ByteBuffer toBuffer()"):src.index("b.rewind();")]
    puts = re.findall(r"\.put(Int|Float)\(", body)
    want = ["Float" if t is C.c_float else "Int" for _, t in abi.ps_config._fields_]
    assert puts == want
    assert "ABI_VERSION = %d" % abi.PS_ABI_VERSION in src
    jni = open(os.path.join(ROOT, "java", "jni", "puckswarm_jni.c")).read()
    native = set(re.findall(r"private static native \w+ (ps\w+)\(", open(os.path.join(ROOT, "java", "puckswarm", "b200", "NativeEngine.java")).read()))
    assert native and all("FN(%s)" % n in jni for n in native)
