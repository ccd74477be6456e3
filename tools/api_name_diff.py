#!/usr/bin/env python
"""API-name parity check: every top-level class / function / constant and every method the reference defines in
``fl4health/<module>.py`` must resolve (by attribute lookup, so inherited methods and re-exports count) in
``fl4health_b200/<module>.py``.  Prints ``<names checked> <names missing>`` and the missing ones.

    python tools/api_name_diff.py
"""
import ast, importlib, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
ref = Path('/root/reference/fl4health')
def names(path):
    out=[]
    tree=ast.parse(path.read_text())
    for node in tree.body:
        if isinstance(node,(ast.FunctionDef,ast.AsyncFunctionDef)): out.append((node.name,None))
        elif isinstance(node,ast.ClassDef):
            out.append((node.name,None))
            for sub in node.body:
                if isinstance(sub,(ast.FunctionDef,ast.AsyncFunctionDef)): out.append((node.name,sub.name))
        elif isinstance(node, ast.Assign):
            for t in node.targets:
                if isinstance(t, ast.Name): out.append((t.id,None))
    return out
total=0; missing=[]
for p in sorted(ref.rglob('*.py')):
    rel=p.relative_to(ref)
    modname='fl4health_b200.'+'.'.join(rel.with_suffix('').parts)
    if modname.endswith('.__init__'): modname=modname[:-9]
    try: mod=importlib.import_module(modname)
    except Exception as e:
        n=names(p); total+=len(n)
        if n: missing.append((str(rel),'MODULE', repr(e)[:80]))
        continue
    for cls,meth in names(p):
        total+=1
        if not hasattr(mod,cls): missing.append((str(rel),cls,meth)); continue
        if meth and not hasattr(getattr(mod,cls),meth): missing.append((str(rel),cls,meth))
print(total,len(missing))
for m in missing: print(*m)
