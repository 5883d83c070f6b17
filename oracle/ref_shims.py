"""TEST INFRASTRUCTURE ONLY -- import shims for running the *reference* in this container.

The reference (`/root/reference/code`) imports packages that are absent from the image
(`matplotlib`, `torchmetrics`, `diffusers`, `accelerate`).  None of them is on the hot path:
`scheduler.py:10` / `sampler.py:13,21` / `trainer_masked.py:7,12` only import them.  This module
registers empty stand-ins so the reference files can be imported on CPU to (a) validate the
restatement in `oracle/mdm_oracle.py` and (b) generate the golden vectors committed under
`tests/golden/` (see `tests/golden/make_golden.py`).

`/root/reference` does not exist on the GPU box: nothing under `tests/ -m gpu`, `smoke()` or
`bench.py` imports this file.
"""
import importlib
import os
import sys
import types

REFERENCE_CODE = os.environ.get("MDM_REFERENCE_CODE", "/root/reference/code")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_CODE, "scheduler.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install_stubs():
    class _Dummy:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

        def __getattr__(self, item):
            return _Dummy()

    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.ticker = _stub("matplotlib.ticker", MaxNLocator=_Dummy)
    try:
        import cv2  # noqa: F401   (tester.py:15 imports it; only its plotting paths use it)
    except Exception:
        _stub("cv2")
    tm = _stub("torchmetrics")
    tm.image = _stub("torchmetrics.image")
    tm.image.fid = _stub("torchmetrics.image.fid", FrechetInceptionDistance=_Dummy)


def import_reference(*names):
    """Return the reference modules `names` (e.g. 'scheduler', 'sampler') imported from
    /root/reference/code under private names so they never shadow this repo's drop-in modules."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    install_stubs()
    out = []
    saved_path = list(sys.path)
    saved = {k: sys.modules.get(k) for k in ("scheduler", "sampler", "trainer_masked",
                                             "trainer_masked_mean_shift", "tester", "utils",
                                             "utils.datautils", "utils.util")}
    try:
        sys.path.insert(0, REFERENCE_CODE)
        for k in saved:
            sys.modules.pop(k, None)
        for n in names:
            out.append(importlib.import_module(n))
    finally:
        # re-home under private names and restore whatever was there before
        for k in list(sys.modules):
            if k in saved or k.startswith("utils."):
                mod = sys.modules.pop(k)
                sys.modules["_mdmref_" + k] = mod
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
        sys.path[:] = saved_path
    return out[0] if len(out) == 1 else out
