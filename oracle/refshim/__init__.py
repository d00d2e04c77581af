"""ORACLE SUPPORT (test infrastructure only -- never imported by the product path).

Loader that runs the reference's OWN source files, unmodified, where they lie under
/root/reference, on the numpy stand-ins for Dr.Jit / Mitsuba in this directory:

    ref = load_reference()          # None when /root/reference is absent (e.g. on the GPU box)
    tree = ref.kdtree.KDTree(); tree.setup(...); tree.addDataPropagate(record); tree.refine() ...

Nothing is copied: the modules are imported from /root/reference/src with `drjit` and `mitsuba`
resolving to oracle/refshim/drjit.py and oracle/refshim/mitsuba.py for the duration of the import
(the previous sys.modules / sys.path entries are restored afterwards, so tests/mitsuba_stub and a
real Mitsuba installation are left alone).  tests/test_reference_on_shim.py uses this leg to pin
oracle/sdtree_oracle.py; tests/golden/make_reference_golden.py freezes its outputs as fixtures.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('SDT_REFERENCE_ROOT', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHADOWED = ('drjit', 'mitsuba', 'scene_stub', 'src', 'src.common', 'src.quadtree', 'src.kdtree',
             'src.path_guiding_integrator', 'src.file_name_manager')
_cache = {}


def available(root=None):
    return os.path.isfile(os.path.join(root or REFERENCE_ROOT, 'src', 'kdtree.py'))


def load_reference(root=None, with_integrator=True):
    """-> namespace(dr, mi, common, quadtree, kdtree, integrator, root) or None"""
    root = root or REFERENCE_ROOT
    if not available(root):
        return None
    key = (root, with_integrator)
    if key in _cache:
        return _cache[key]
    saved_modules = {k: sys.modules.pop(k) for k in _SHADOWED if k in sys.modules}
    saved_path = list(sys.path)
    repo_root = os.path.dirname(os.path.dirname(_HERE))
    dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True            # /root/reference is read-only
    try:
        sys.path[:0] = [_HERE, root, repo_root]
        ns = types.SimpleNamespace(root=root)
        ns.dr = importlib.import_module('drjit')
        ns.mi = importlib.import_module('mitsuba')
        ns.mi.set_variant('shim_numpy_rgb')
        ns.common = importlib.import_module('src.common')
        ns.quadtree = importlib.import_module('src.quadtree')
        ns.kdtree = importlib.import_module('src.kdtree')
        ns.integrator = importlib.import_module('src.path_guiding_integrator') if with_integrator else None
        for m in (ns.common, ns.quadtree, ns.kdtree, ns.integrator):
            assert m is None or os.path.realpath(m.__file__).startswith(os.path.realpath(root)), m
    finally:
        sys.dont_write_bytecode = dont_write
        sys.path[:] = saved_path
        for k in _SHADOWED:
            sys.modules.pop(k, None)
        sys.modules.update(saved_modules)
    _cache[key] = ns
    return ns


def run_reference_main(name, root=None):
    """executes the `if __name__ == '__main__':` self-test of /root/reference/src/<name>.py, unmodified, on
    the stand-ins -> (module globals, captured stdout)"""
    import contextlib
    import io
    import runpy
    ref = load_reference(root)
    saved = {k: sys.modules.get(k) for k in _SHADOWED}
    src_pkg = types.ModuleType('src')
    src_pkg.__path__ = [os.path.join(ref.root, 'src')]
    sys.modules.update({'drjit': ref.dr, 'mitsuba': ref.mi, 'src': src_pkg, 'src.common': ref.common,
                        'src.quadtree': ref.quadtree, 'src.kdtree': ref.kdtree})
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            g = runpy.run_path(os.path.join(ref.root, 'src', name + '.py'), run_name='__main__')
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return g, buf.getvalue()
