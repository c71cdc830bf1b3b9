#!/usr/bin/env python
"""Build the UNMODIFIED reference (hydpy 6.5dev0) hot path as binaries.  TEST INFRASTRUCTURE ONLY.

The reference is Python + generated Cython.  This recipe

1. copies ``/root/reference`` to a scratch directory outside the repository (the reference tree is
   read-only and its build writes into ``hydpy/cythons/autogen``),
2. puts tiny stub modules for the four import-time-only dependencies that are absent in this image
   (``black``, ``inflect``, ``netCDF4``, ``matplotlib``; none touches the arithmetic) on ``PYTHONPATH``,
3. runs the reference's *own* generator (``prepare_build.py`` steps: base ``*utils.pyx``, the
   ``masterinterface``, and ``Cythonizer.pyxwriter.write()`` for the nine hot-path model modules) and
   compiles the generated sources with the reference's own flags (``-O2``; ``cdivision``/``cpow`` etc.
   come from the directives the generator writes into each ``.pyx``),
4. copies the resulting shared objects into ``oracle/_ref/hydpy/cythons/autogen/`` and — like ``pip install --target``
   would — the reference's Python package beside them (``oracle/_ref/hydpy``, plus the stubs in ``oracle/_ref/_stubs``),
   so that ``oracle/_ref`` is an importable installation of the unmodified reference.  The directory is git-ignored,
   not gpurun-ignored: it travels to the GPU box, where the ``-m gpu`` drop-in tests run ``HydPy.simulate`` of real
   reference objects against the engine.  No reference file enters the repository's history.

The scratch copy (default ``/tmp/hydpy_b200_refbuild/src``) stays importable in THIS container and is what
``tests/golden/make_golden.py`` uses to generate the committed fixtures.

Usage:  python oracle/build_ref.py [--scratch DIR] [--reference /root/reference] [--force]
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "hydpy", "cythons", "autogen")
DEFAULT_SCRATCH = os.environ.get("HYDPY_B200_REFBUILD", "/tmp/hydpy_b200_refbuild")

MODELS = (
    "hland", "hland_96", "evap", "evap_aet_hbv96", "evap_pet_hbv96",
    "rconc", "rconc_uh", "musk", "musk_classic",
    # sibling models on the same skeleton
    "hland_96p", "hland_96c", "rconc_nash", "musk_mct", "wq", "wq_trapeze_strickler",
    # the second PET submodel behind evap_aet_hbv96 (external reference evapotranspiration)
    "evap_ret_io",
)

STUBS = {
    "black.py": """
        class FileMode:
            def __init__(self, *args, **kwargs):
                pass
        Mode = FileMode
        def format_str(string, mode=None, **kwargs):
            return string
    """,
    "inflect.py": """
        class engine:
            def plural(self, word, count=None):
                return word if count == 1 else word + "s"
            def plural_noun(self, word, count=None):
                return self.plural(word, count)
            def plural_verb(self, word, count=None):
                return word
            def number_to_words(self, number):
                return str(number)
            def a(self, word):
                return "a " + word
            an = a
    """,
    "netCDF4.py": """
        class Dataset:
            def __init__(self, *args, **kwargs):
                raise ImportError("netCDF4 is stubbed out in the hydpy_b200 oracle build")
    """,
    "matplotlib/__init__.py": """
        def use(*args, **kwargs):
            pass
    """,
    "matplotlib/figure.py": """
        class Figure:
            pass
    """,
    "matplotlib/pyplot.py": """
        def __getattr__(name):
            raise ImportError("matplotlib is stubbed out in the hydpy_b200 oracle build")
    """,
}

BUILD_SNIPPET = r"""
import importlib, os, sys
sys.setrecursionlimit(10000)
import numpy, setuptools, Cython.Build
import prepare_build as pb

autogen = os.path.join("hydpy", "cythons", "autogen")
stage = sys.argv[1]
if stage == "base":
    pb._clear_autogendir()
    pb._prepare_baseextensions(fast_cython=True, profile_cython=False)
    pb._compile_extensions(filetype="utils")
elif stage == "interface":
    pb._convert_interfaces(fast_cython=True, profile_cython=False)
    pb._compile_extensions(filetype="interface")
elif stage == "models":
    from hydpy import config, pub
    config.FASTCYTHON = True
    config.PROFILECYTHON = False
    names = sys.argv[2:]
    with pub.options.usecython(False):
        for name in names:
            module = importlib.import_module(f"hydpy.models.{name}")
            module.cythonizer.pyxwriter.write()
    exts = [
        setuptools.Extension(
            name=f"hydpy.cythons.autogen.c_{name}",
            sources=[os.path.join(autogen, f"c_{name}.pyx")],
            extra_compile_args=["-O2"],
        )
        for name in names
    ]
    sys.argv = [sys.argv[0], "build_ext", "--build-lib=.", "--build-temp=build_tmp", "-j", str(os.cpu_count() or 1)]
    setuptools.setup(
        name="temporary",
        ext_modules=Cython.Build.cythonize(exts, nthreads=os.cpu_count() or 1),
        include_dirs=[numpy.get_include()],
    )
"""


def _run(cmd, cwd, env):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, cwd=cwd, env=env, check=True)


def scratch_paths(scratch: str = DEFAULT_SCRATCH) -> tuple[str, str]:
    """Return (source copy, stubs) directories of a finished scratch build."""
    return os.path.join(scratch, "src"), os.path.join(scratch, "stubs")


INSTALL = os.path.join(HERE, "_ref")
INSTALL_STUBS = os.path.join(INSTALL, "_stubs")


def installed() -> bool:
    """True when oracle/_ref holds the importable installation (Python package + binaries), here or on the GPU box."""
    return os.path.exists(os.path.join(INSTALL, "hydpy", "__init__.py")) and os.path.isdir(INSTALL_STUBS) and any(
        n.startswith("c_hland_96.") and n.endswith(".so") for n in os.listdir(OUT))


def import_paths() -> tuple[str, str] | None:
    """(package root, stubs) to put in front of sys.path for `import hydpy`: the installation in oracle/_ref when it is
    complete, else the scratch build of this container, else None."""
    if installed():
        return INSTALL, INSTALL_STUBS
    src, stubs = scratch_paths()
    if os.path.exists(os.path.join(os.path.dirname(src), ".built")):
        return src, stubs
    return None


def install_package(src: str, stubs: str) -> None:
    """Copy the reference's Python package (no generated C, no build leftovers) and the stubs next to the binaries."""
    skip = shutil.ignore_patterns("*.c", "*.cpp", "*.pyx", "*.html", "*.so", "*.o", "__pycache__", "build", "build_tmp",
                                  "_build", "html_")
    pkg = os.path.join(INSTALL, "hydpy")
    binaries = {n: os.path.join(OUT, n) for n in os.listdir(OUT) if n.endswith(".so")} if os.path.isdir(OUT) else {}
    tmp = os.path.join(INSTALL, "_hydpy_new")
    if os.path.isdir(tmp):
        shutil.rmtree(tmp)
    shutil.copytree(os.path.join(src, "hydpy"), tmp, ignore=skip, symlinks=False)
    os.makedirs(os.path.join(tmp, "cythons", "autogen"), exist_ok=True)
    for name, path in binaries.items():
        shutil.copy2(path, os.path.join(tmp, "cythons", "autogen", name))
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.rename(tmp, pkg)
    if os.path.isdir(INSTALL_STUBS):
        shutil.rmtree(INSTALL_STUBS)
    shutil.copytree(stubs, INSTALL_STUBS, ignore=shutil.ignore_patterns("__pycache__"))


def build(reference: str = "/root/reference", scratch: str = DEFAULT_SCRATCH, force: bool = False) -> str:
    src, stubs = scratch_paths(scratch)
    marker = os.path.join(scratch, ".built")
    if force and os.path.isdir(scratch):
        shutil.rmtree(scratch)
    if not os.path.exists(marker):
        if not os.path.isdir(reference):
            raise FileNotFoundError(f"reference tree {reference!r} is not available; cannot build oracle/_ref")
        os.makedirs(scratch, exist_ok=True)
        if os.path.isdir(src):
            shutil.rmtree(src)
        shutil.copytree(reference, src, symlinks=True)
        for rel, text in STUBS.items():
            path = os.path.join(stubs, rel)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w", encoding="utf-8") as f:
                f.write(textwrap.dedent(text))
        with open(os.path.join(src, "_b200_build.py"), "w", encoding="utf-8") as f:
            f.write(BUILD_SNIPPET)
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([stubs, src, env.get("PYTHONPATH", "")])
        env.setdefault("CC", "gcc")
        for stage in (["base"], ["interface"], ["models", *MODELS]):
            _run([sys.executable, "_b200_build.py", *stage], cwd=src, env=env)
        with open(marker, "w") as f:
            f.write("\n".join(MODELS) + "\n")
    else:
        # a scratch build of an earlier model list: generate and compile only what is missing
        built = os.listdir(os.path.join(src, "hydpy", "cythons", "autogen"))
        missing = [m for m in MODELS if not any(n.startswith(f"c_{m}.") and n.endswith(".so") for n in built)]
        if missing:
            env = dict(os.environ)
            env["PYTHONPATH"] = os.pathsep.join([stubs, src, env.get("PYTHONPATH", "")])
            env.setdefault("CC", "gcc")
            _run([sys.executable, "_b200_build.py", "models", *missing], cwd=src, env=env)
            with open(marker, "w") as f:
                f.write("\n".join(MODELS) + "\n")
    # copy the binaries (and nothing else) into oracle/_ref
    os.makedirs(OUT, exist_ok=True)
    autogen = os.path.join(src, "hydpy", "cythons", "autogen")
    copied = []
    for name in sorted(os.listdir(autogen)):
        if name.endswith(".so"):
            shutil.copy2(os.path.join(autogen, name), os.path.join(OUT, name))
            copied.append(name)
    print(f"oracle/_ref: {len(copied)} shared objects -> {OUT}")
    install_package(src, stubs)
    print(f"oracle/_ref: Python package of the reference + stubs installed beside them ({INSTALL})")
    return src


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--scratch", default=DEFAULT_SCRATCH)
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    build(args.reference, args.scratch, args.force)


if __name__ == "__main__":
    main()
