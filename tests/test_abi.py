"""The C-ABI boundary without a GPU: `include/hydpy_b200.h` is the single source of truth — every function it declares
must be exported by the built library and bound by the ctypes layer, every enum/define must agree with
`hydpy_b200/layout.py`, and without a CUDA device the library must refuse to work (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from hydpy_b200 import abi, engine, layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "hydpy_b200.h")).read()


def declared_functions() -> list[str]:
    code = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return re.findall(r"^\s*(?:const\s+char\s*\*|int|void)\s+(hb_\w+)\s*\(", code, flags=re.M)


def enum_members(name: str) -> list[str]:
    body = re.search(r"enum\s+%s\s*\{(.*?)\}" % name, HEADER, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [m.split("=")[0].strip() for m in body.split(",") if m.strip()]


def define(name: str) -> int:
    m = re.search(r"#define\s+%s\s+\(?(-?\(?[\w<\- ]+\)?)\)?" % name, HEADER)
    text = m.group(1).strip()
    return int(eval(text.replace("(", "").replace(")", ""), {})) if "<<" not in text else -(1 << 30)


def test_header_declares_what_python_binds():
    assert sorted(declared_functions()) == sorted(engine.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(engine.LIBPATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/hydpy_b200.h but not exported"
    lib.hb_abi_version.restype = C.c_int
    assert lib.hb_abi_version() == define("HB_ABI_VERSION") == layout.ABI_VERSION


def test_enums_match_layout():
    assert [m[len("HB_ZP_"):].lower() for m in enum_members("hb_zpar")[:-1]] == list(layout.ZPAR)
    assert [m[len("HB_EP_"):].lower() for m in enum_members("hb_epar")[:-1]] == list(layout.EPAR)
    assert [m[len("HB_F_"):].lower() for m in enum_members("hb_forcing")[:-1]] == list(layout.FORCING)
    series = [m[len("HB_S_"):].lower() for m in enum_members("hb_series")[:-1]]
    assert series == [s.rstrip("_") for s in layout.SERIES]
    assert [m[len("HB_MS_"):].lower() for m in enum_members("hb_mct_series")[:-1]] == list(layout.MCT_SERIES)
    for cname, value in (("HB_FIELD", layout.FIELD), ("HB_FOREST", layout.FOREST), ("HB_GLACIER", layout.GLACIER),
                         ("HB_ILAKE", layout.ILAKE), ("HB_SEALED", layout.SEALED), ("HB_ZF_WATER", layout.ZF_WATER),
                         ("HB_ZF_INTERCEPTION", layout.ZF_INTERCEPTION), ("HB_ZF_SOIL", layout.ZF_SOIL),
                         ("HB_ZF_TREE", layout.ZF_TREE), ("HB_ZF_SREDEND", layout.ZF_SREDEND),
                         ("HB_EF_RESPAREA", layout.EF_RESPAREA), ("HB_NODE_NEWSIM", layout.NODE_NEWSIM),
                         ("HB_NODE_EXT", layout.NODE_EXT), ("HB_NODE_OBSNEWSIM", layout.NODE_OBSNEWSIM),
                         ("HB_RUN_LAND", layout.RUN_LAND), ("HB_RUN_RIVER", layout.RUN_RIVER),
                         ("HB_ROW_NODE", layout.ROW_NODE), ("HB_ROW_REACH", layout.ROW_REACH),
                         ("HB_OPT_LAND_PATH", layout.OPT_LAND_PATH), ("HB_OPT_RIVER_SWEEP", layout.OPT_RIVER_SWEEP), ("HB_OPT_RIVER_STATE_SERIES", layout.OPT_RIVER_STATE_SERIES),
                         ("HB_OPT_RIVER_WAVE_BLOCKS", layout.OPT_RIVER_WAVE_BLOCKS),
                         ("HB_OPT_RIVER_MCT_SERIES", layout.OPT_RIVER_MCT_SERIES),
                         ("HB_OPT_LAND_SLABS", layout.OPT_LAND_SLABS), ("HB_MOD_EPAR", layout.MOD_EPAR), ("HB_OPT_LAND_CHUNK", layout.OPT_LAND_CHUNK),
                         ("HB_RIVER_PERSISTENT", layout.RIVER_PERSISTENT), ("HB_RIVER_LAUNCHES", layout.RIVER_LAUNCHES), ("HB_RIVER_GROUPED", layout.RIVER_GROUPED), ("HB_LAND_AUTO", layout.LAND_AUTO),
                         ("HB_LAND_GENERAL", layout.LAND_GENERAL), ("HB_NCCL_ID_BYTES", layout.NCCL_ID_BYTES),
                         ("HB_ENTRY_EXT", layout.ENTRY_EXT)):
        assert define(cname) == value, cname


def test_struct_fields_follow_the_header():
    for cstruct, pystruct in (("hb_land_desc", abi.hb_land_desc), ("hb_land_states", abi.hb_land_states),
                              ("hb_river_desc", abi.hb_river_desc), ("hb_timing", engine.hb_timing)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cstruct, cstruct), HEADER, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [re.findall(r"(\w+)\s*$", decl.strip())[0] for decl in body.split(";") if decl.strip()]
        assert names == [f[0] for f in pystruct._fields_], cstruct


@pytest.mark.skipif(engine.device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(engine.EngineUnavailable):
        engine.Engine(0)
