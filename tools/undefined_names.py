"""Tiny stand-in for pyflakes' undefined-name check (no linter in this image): reports names that are loaded in a
function but bound nowhere visible (locals, enclosing functions, module globals, builtins).
usage: python tools/undefined_names.py file.py [...]"""
import ast
import builtins
import sys


def bound_names(node):
    out = set()
    for n in ast.walk(node):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        if isinstance(n, ast.arg):
            out.add(n.arg)
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                out.add((a.asname or a.name).split(".")[0])
        if isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        if isinstance(n, (ast.Global, ast.Nonlocal)):
            out.update(n.names)
    return out


def check(path):
    tree = ast.parse(open(path).read(), path)
    module_names = bound_names(tree) | set(dir(builtins)) | {"__file__", "__name__"}
    bad = []

    def visit(fn, enclosing):
        scope = enclosing | bound_names(fn)
        for n in ast.walk(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in scope:
                bad.append((path, n.lineno, n.id))

    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef)):
            visit(n, module_names)
    return bad


if __name__ == "__main__":
    problems = [b for p in sys.argv[1:] for b in check(p)]
    for p, line, name in problems:
        print(f"{p}:{line}: undefined name {name!r}")
    sys.exit(1 if problems else 0)
