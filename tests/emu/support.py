"""CPU-side test support: builds the thread-emulated kernel library for a standard model and returns a COCSys whose
memory provider is numpy and whose library is the emulation.  TEST INFRASTRUCTURE — the product never imports this."""
import hashlib
import os

import numpy as np

import lfsd_b200  # noqa: F401
from lfsd_b200 import _capi, codegen, standard
from tests.emu.build_emu import build

_BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")


class NumpyMem:
    def empty(self, shape, dtype="f8"):
        return np.zeros(shape, dtype={"f8": np.float64, "i4": np.int32, "u1": np.uint8}[dtype])

    zeros = empty

    def from_host(self, a, dtype="f8"):
        return np.array(a, dtype={"f8": np.float64, "i4": np.int32}[dtype], order="C", copy=True)

    def to_host(self, x):
        return np.asarray(x)

    def ptr(self, x):
        return 0 if x is None else x.ctypes.data

    def stream(self):
        return 0


def emu_oc(name, tsan=False, **kw):
    """COCSys for standard model `name` bound to the host-emulated kernels."""
    oc = (standard.STANDARD.get(name) or standard.VARIANTS[name])(**kw)
    text, info = codegen.generate_model_header(name, oc.state, oc.control, oc.auxvar, oc.dyn, oc.path_cost,
                                               oc.final_cost, oc.pvar, time=getattr(oc, 'time', None))
    os.makedirs(_BUILD, exist_ok=True)
    src = ""
    for fn in sorted(os.listdir(_capi.CSRC)):
        p = os.path.join(_capi.CSRC, fn)
        if os.path.isfile(p):
            src += open(p).read()
    src += open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu_lib.cpp")).read()
    tag = hashlib.sha1((text + src + str(tsan)).encode()).hexdigest()[:12]
    hdr = os.path.join(_BUILD, "model_%s_%s.cuh" % (name, tag))
    so = os.path.join(_BUILD, "libemu_%s_%s.so" % (name, tag))
    if not os.path.exists(so):
        with open(hdr, "w") as f:
            f.write(text)
        build(hdr, so, tsan=tsan, ns="cpdp_emu_" + name, extra=["-DCPDP_WITH_BDF"] if os.path.exists(os.path.join(_capi.CSRC, "cpdp_bdf.cuh")) else [])
    oc._lib = _capi.CpdpLib(so)
    oc._mem = NumpyMem()
    return oc
