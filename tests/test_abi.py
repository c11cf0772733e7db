"""CPU: the C-ABI library loads and exports every symbol include/bbmap_cuda.h declares plus the reference's JNI
symbol names; without a device the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "bbmap_b200", "libbbmapcuda.so")

JNI_SYMBOLS = [
    # jni/align2_MultiStateAligner11tsJNI.h:165-174
    "Java_align2_MultiStateAligner11tsJNI_fillUnlimitedJNI",
    "Java_align2_MultiStateAligner11tsJNI_fillLimitedXJNI",
    # jni/align2_BandedAlignerJNI.h:17-41
    "Java_align2_BandedAlignerJNI_alignForwardJNI",
    "Java_align2_BandedAlignerJNI_alignForwardRCJNI",
    "Java_align2_BandedAlignerJNI_alignReverseJNI",
    "Java_align2_BandedAlignerJNI_alignReverseRCJNI",
]


def _declared():
    src = open(os.path.join(ROOT, "include", "bbmap_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bbm_[a-zA-Z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(SO):
        g.build()
    return C.CDLL(SO)


def test_exports(lib):
    names = _declared()
    assert len(names) >= 12
    for n in names + JNI_SYMBOLS:
        assert hasattr(lib, n), "libbbmapcuda.so does not export %s" % n
    from bbmap_b200 import lib as L
    assert set(L.EXPORTS) == set(names), sorted(set(names) ^ set(L.EXPORTS))      # every declared entry point has ctypes argtypes


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib.bbm_last_error.restype = C.c_char_p
    h = C.c_void_p()
    assert lib.bbm_device_count() == 0
    assert lib.bbm_init(0, C.byref(h)) == -1          # BBM_E_NODEVICE
    assert b"no CPU fallback" in lib.bbm_last_error()
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    from bbmap_b200.lib import BbmError
    with pytest.raises(BbmError):
        MultiStateAligner11tsCUDA()


def test_record_layouts():
    from bbmap_b200 import workloads as wl
    assert wl.TASK_DTYPE.itemsize == 40 and wl.OUT_DTYPE.itemsize == 80
    assert wl.OUT_DTYPE.fields["iterations"][1] == 24 and wl.OUT_DTYPE.fields["score"][1] == 32
