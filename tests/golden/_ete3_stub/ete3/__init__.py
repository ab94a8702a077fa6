"""Stand-in for the absent ``ete3`` package, used ONLY by tests/golden/make_golden.py so the
unmodified reference module (/root/reference/src/ancestral_reconstruction.py:2) imports in
this container.  It re-exports the package's own ete3-free tree; ``ete3.treeview`` is
deliberately missing so the reference's plot step degrades gracefully (AR:821-826)."""
import importlib.util as _u
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg = _os.path.abspath(_os.path.join(_here, "..", "..", "..", "..", "chromevol_enhanced_b200", "newick.py"))
_spec = _u.spec_from_file_location("_cevb_newick_for_stub", _pkg)
_mod = _u.module_from_spec(_spec)
_sys.modules[_spec.name] = _mod
_spec.loader.exec_module(_mod)
Tree = _mod.Tree
TreeNode = _mod.TreeNode
