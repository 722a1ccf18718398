"""oracle/reference.py -- TEST / BASELINE INFRASTRUCTURE ONLY (never imported by mass_b200/).

Loads the UNMODIFIED reference (brandontrabucco/mass) so that it can be (1) the differential check of the C
oracle (tests/test_oracle_vs_reference.py), (2) the generator of tests/golden/*.npz, and (3) the CPU arm of
bench.py (`--impl reference`, `cpu_baseline.kind == "reference"`): the reference's own torch CPU path,
mass/nn/base_projection_layer.py:282-343, timed on the box's host cores.

Where it comes from: `install()` (run by __graft_entry__.build() wherever /root/reference exists) pip-installs the
reference package, from a scratch copy of its tree, into the git-ignored directory oracle/_ref/ -- no reference
source is ever committed.  oracle/_ref/ travels to the GPU box with the working-tree snapshot, /root/reference does
not; `root()` prefers the source tree when present, else the installed copy.

Shims: the applications import a stale package name ``slam_rcnn`` -> aliased to ``mass``;
experimentation.py imports simulator packages that are not installed -> three stub modules holding only the
names it reads (the class tables restate segmentation_config.py:43-117: id 0 neither, 1-43 pickable, 44-53 openable).
"""
import os
import shutil
import subprocess
import sys
import tempfile
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE_ROOT = os.environ.get("MASS_REFERENCE_ROOT", "/root/reference")
INSTALL_ROOT = os.path.join(_HERE, "_ref")


def root():
    """Directory to put on sys.path: the reference tree if this machine has it, else the installed copy."""
    if os.path.isdir(os.path.join(SOURCE_ROOT, "mass")):
        return SOURCE_ROOT
    if os.path.isdir(os.path.join(INSTALL_ROOT, "mass")):
        return INSTALL_ROOT
    return None


def available():
    return root() is not None


def install(force=False):
    """pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>.
    The copy is needed because the build writes into the source tree and /root/reference is read-only.  Returns
    True if oracle/_ref holds the package afterwards."""
    have = os.path.isdir(os.path.join(INSTALL_ROOT, "mass"))
    if not os.path.isdir(os.path.join(SOURCE_ROOT, "mass")):
        return have
    if have and not force:
        return True
    tmp = tempfile.mkdtemp(prefix="mass_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(SOURCE_ROOT, src, ignore=shutil.ignore_patterns("*.pth", "images", ".git"))
        shutil.rmtree(INSTALL_ROOT, ignore_errors=True)
        proc = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                               "--find-links", "/opt/wheelhouse", "--target", INSTALL_ROOT, src],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("installing the reference into oracle/_ref failed:\n" + proc.stdout[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return os.path.isdir(os.path.join(INSTALL_ROOT, "mass"))


def load():
    """Returns a namespace with the reference's modules."""
    where = root()
    if where is None:
        raise ImportError("reference not present at %s nor installed in %s" % (SOURCE_ROOT, INSTALL_ROOT))
    warnings.filterwarnings("ignore", category=UserWarning)
    if where not in sys.path:
        sys.path.insert(0, where)
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
