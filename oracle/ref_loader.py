"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules from
/root/reference (present in the build container, absent on the GPU box).

Used by oracle/make_golden.py and by the CPU tests that pin the oracle against
the reference itself.  Nothing in das4whales_b200/ may import this.

The reference imports librosa / sparse / matplotlib at module top
(/root/reference/src/das4whales/dsp.py:12-13, detect.py:10-16); those are not
installed here, so empty stand-ins are injected before loading dsp.py / plot.py /
detect.py by path.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("D4W_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REF_ROOT, "src", "das4whales")


def available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "dsp.py"))


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


_cache = {}


def load():
    """Return (dsp, detect) reference modules (unmodified source)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise FileNotFoundError(f"reference not found under {REF_ROOT}")
    for n in ("librosa", "sparse", "matplotlib", "matplotlib.pyplot", "matplotlib.ticker",
              "matplotlib.colors", "matplotlib.gridspec", "tqdm"):
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    sp_mod = sys.modules["sparse"]
    if not hasattr(sp_mod, "COO"):
        sp_mod.COO = type("COO", (), {"from_numpy": staticmethod(lambda a: a)})
    tq = sys.modules["tqdm"]
    if not hasattr(tq, "tqdm"):
        tq.tqdm = lambda it, **kw: it
    pkg = types.ModuleType("das4whales")
    pkg.__path__ = [_SRC]
    sys.modules.setdefault("das4whales", pkg)

    def _load(name):
        spec = importlib.util.spec_from_file_location(f"das4whales.{name}", os.path.join(_SRC, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        return m

    dsp = _load("dsp")
    try:
        _load("plot")
    except Exception:
        # plot.py needs a real matplotlib; detect.py only wants import_roseus from it
        p = types.ModuleType("das4whales.plot")
        p.import_roseus = lambda: None
        sys.modules["das4whales.plot"] = p
    detect = _load("detect")
    _cache["mods"] = (dsp, detect)
    return dsp, detect


def load_data_handle():
    """The reference's data_handle module (unmodified) with its file-format dependencies stubbed out; only the
    arithmetic helper raw2strain (data_handle.py:157-177) is used from it."""
    if "dh" in _cache:
        return _cache["dh"]
    load()
    for n in ("h5py", "wget", "nptdms", "dask", "dask.array", "xarray", "pyproj", "pandas"):
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    if not hasattr(sys.modules["nptdms"], "TdmsFile"):
        sys.modules["nptdms"].TdmsFile = type("TdmsFile", (), {})
    if "dask" in sys.modules and not hasattr(sys.modules["dask"], "array") and "dask.array" in sys.modules:
        sys.modules["dask"].array = sys.modules["dask.array"]
    spec = importlib.util.spec_from_file_location("das4whales.data_handle", os.path.join(_SRC, "data_handle.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = m
    spec.loader.exec_module(m)
    _cache["dh"] = m
    return m


def load_improcess():
    """The reference's improcess module (unmodified).  Its real dependencies cv2 / torch / torchvision are importable here;
    skimage and matplotlib are stubbed (only used by functions outside the Gabor path)."""
    if "imp" in _cache:
        return _cache["imp"]
    load()
    for n in ("skimage", "skimage.transform"):
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    st = sys.modules["skimage.transform"]
    for name in ("radon", "iradon"):
        if not hasattr(st, name):
            setattr(st, name, lambda *a, **k: None)
    sys.modules["skimage"].transform = st
    spec = importlib.util.spec_from_file_location("das4whales.improcess", os.path.join(_SRC, "improcess.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = m
    spec.loader.exec_module(m)
    _cache["imp"] = m
    return m
