"""Imports the UNMODIFIED reference (/root/reference) for golden-vector generation
and differential oracle tests.  Only usable in the build container: the GPU box
has no /root/reference, and nothing on the `-m gpu` path may import this module.

Shims: the applications import a stale package name
``slam_rcnn`` -> aliased to ``mass``; experimentation.py imports simulator
packages that are not installed -> three stub modules holding only the names it
reads (the class tables restate segmentation_config.py:43-117: id 0 neither,
1-43 pickable, 44-53 openable).
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("MASS_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mass"))


def load():
    """Returns a namespace with the reference's modules."""
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=UserWarning)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mass
    import mass.nn
    import mass.nn.base_projection_layer as base
    import mass.utils
    import mass.utils.projection as projection

    # stale package alias
    sys.modules.setdefault("slam_rcnn", mass)
    sys.modules.setdefault("slam_rcnn.nn", mass.nn)
    sys.modules.setdefault("slam_rcnn.nn.base_projection_layer", base)
    sys.modules.setdefault("slam_rcnn.utils", mass.utils)
    viz = types.ModuleType("slam_rcnn.utils.visualization")
    viz.visualize_path = lambda *a, **k: None
    sys.modules.setdefault("slam_rcnn.utils.visualization", viz)

    # simulator stubs for experimentation.py
    if "rearrange.tasks" not in sys.modules:
        rearrange = types.ModuleType("rearrange")
        tasks = types.ModuleType("rearrange.tasks")
        tasks.UnshuffleTask = type("UnshuffleTask", (), {})
        rearrange.tasks = tasks
        sys.modules["rearrange"] = rearrange
        sys.modules["rearrange.tasks"] = tasks
    if "ai2thor.exceptions" not in sys.modules:
        ai2thor = types.ModuleType("ai2thor")
        exc = types.ModuleType("ai2thor.exceptions")
        exc.RestartError = type("RestartError", (Exception,), {})
        exc.UnityCrashException = type("UnityCrashException", (Exception,), {})
        ai2thor.exceptions = exc
        sys.modules["ai2thor"] = ai2thor
        sys.modules["ai2thor.exceptions"] = exc
    if "mass.thor.segmentation_config" not in sys.modules:
        thor = types.ModuleType("mass.thor")
        seg = types.ModuleType("mass.thor.segmentation_config")
        names = ["OccupiedSpace"] + ["pickable%d" % i for i in range(43)] + \
                ["openable%d" % i for i in range(10)]
        seg.PICKABLE_TO_COLOR = {n: (0, 0, 0) for n in names[1:44]}
        seg.OPENABLE_TO_COLOR = {n: (0, 0, 0) for n in names[44:]}
        seg.ID_TO_PICKABLE = [n in seg.PICKABLE_TO_COLOR for n in names]
        seg.ID_TO_OPENABLE = [n in seg.OPENABLE_TO_COLOR for n in names]
        thor.segmentation_config = seg
        sys.modules["mass.thor"] = thor
        sys.modules["mass.thor.segmentation_config"] = seg

    import mass.nn.applications.semantic_projection_layer as semantic
    import mass.nn.applications.occupancy_projection_layer as occupancy
    import mass.utils.experimentation as experimentation
    return types.SimpleNamespace(projection=projection, base=base, semantic=semantic,
                                 occupancy=occupancy, experimentation=experimentation)
