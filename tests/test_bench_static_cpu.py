"""bench.py and tools/ only run on the GPU box; this catches unbound names in them on the CPU box (a poor man's pyflakes:
every name a function loads must be bound in that function, an enclosing one, the module, or builtins)."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["bench.py", "__graft_entry__.py"] + [os.path.join("tools", f) for f in sorted(os.listdir(os.path.join(ROOT, "tools")))
                                             if f.endswith(".py")]


def _walk_shallow(node):
    """ast.walk that does not descend into nested function bodies (their locals are not the enclosing scope's)."""
    todo = list(ast.iter_child_nodes(node))
    while todo:
        n = todo.pop()
        yield n
        if not isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            todo.extend(ast.iter_child_nodes(n))


def _bound(node, shallow=False):
    names = set()
    for n in (_walk_shallow(node) if shallow else ast.walk(node)):
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            names.add(n.id)
        elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(n.name)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(n, ast.arg):
            names.add(n.arg)
        elif isinstance(n, ast.ExceptHandler) and n.name:
            names.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            names.update(n.names)
    return names


@pytest.mark.parametrize("rel", FILES)
def test_no_unbound_names(rel):
    tree = ast.parse(open(os.path.join(ROOT, rel)).read())
    module_names = _bound(tree, shallow=True) | set(dir(builtins)) | {"__file__", "__name__"}
    problems = []

    def visit(fn, outer):
        scope = outer | _bound(fn, shallow=True)
        for a in ast.walk(fn.args):
            if isinstance(a, ast.arg):
                scope.add(a.arg)
        first_store = {}
        comp_vars = {t.id for n in _walk_shallow(fn) if isinstance(n, ast.comprehension) for t in ast.walk(n.target)
                     if isinstance(t, ast.Name)}
        for n in _walk_shallow(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store) and n.id not in comp_vars:
                first_store[n.id] = min(first_store.get(n.id, 1 << 30), n.lineno)
        for n in _walk_shallow(fn):
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in scope:
                problems.append(f"{rel}:{n.lineno}: {n.id} in {getattr(fn, 'name', '<lambda>')}")
            elif isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in outer and \
                    n.id in first_store and n.lineno < first_store[n.id]:
                # a local read on a line above its first assignment (straight-line benchmark code: no such loops here)
                problems.append(f"{rel}:{n.lineno}: {n.id} read before assignment in {getattr(fn, 'name', '<lambda>')}")
            elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                visit(n, scope)

    for node in _walk_shallow(tree):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            visit(node, module_names)
    assert not problems, problems
