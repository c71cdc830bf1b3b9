#!/usr/bin/env python
"""Poor man's pyflakes (not installed here): names that a function loads but nothing in its scopes, the module or the
builtins defines.  python tools/undefined_names.py file.py [...]"""
import ast
import builtins
import sys


def check(path: str) -> int:
    tree = ast.parse(open(path).read(), path)
    module_names = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            module_names.update((a.asname or a.name).split(".")[0] for a in node.names)
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            module_names.add(node.name)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            module_names.add(node.id)
        elif isinstance(node, ast.arg):
            module_names.add(node.arg)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            module_names.add(node.name)
    bad = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load) and node.id not in module_names:
            print(f"{path}:{node.lineno}: undefined name {node.id!r}")
            bad += 1
    return bad


if __name__ == "__main__":
    sys.exit(1 if sum(check(p) for p in sys.argv[1:]) else 0)
