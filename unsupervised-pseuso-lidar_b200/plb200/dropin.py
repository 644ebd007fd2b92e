"""Makes the reference's own import statements resolve to the B200 modules.

    import plb200.dropin; plb200.dropin.install()        # first lines of train.py / PseudoLidarPipeline.py

`sys.path` shadowing alone is enough for `losses`, `geometry.pose_geometry`, `geometry.transform` (the reference's
`geometry/` has no `__init__.py`, so the two directories merge into one namespace package and the first path
entry wins per module) and for `Transform.Transform`.  It is NOT enough for `utils.PseudoLiDAR`: the reference's
`pseudo-lidar/utils/` has an `__init__.py`, a regular package always beats a namespace portion, and giving our
`utils/` an `__init__.py` would in turn hide the reference's `utils.model` / `utils.transforms`.  So the two
pseudo-LiDAR modules are registered in `sys.modules` under the names the reference imports."""
import importlib
import os
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(pseudo_lidar=True):
    """Idempotent.  Returns the list of module names that now resolve to this repository."""
    if PKG_DIR in sys.path:
        sys.path.remove(PKG_DIR)
    sys.path.insert(0, PKG_DIR)
    names = ["losses", "geometry.pose_geometry", "geometry.transform"]
    if pseudo_lidar:
        from . import pseudolidar, velodyne
        for pkg_name, mod_name, mod in (("utils", "PseudoLiDAR", pseudolidar), ("Transform", "Transform", velodyne)):
            full = pkg_name + "." + mod_name
            sys.modules[full] = mod
            try:
                pkg = importlib.import_module(pkg_name)      # the reference's package when it is the one found
            except ImportError:
                continue
            setattr(pkg, mod_name, mod)
            names.append(full)
    return names
