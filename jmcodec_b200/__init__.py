"""jmcodec_b200 -- B200-native decoded-surface format path behind jmcodec's jm_nvdec_* / jm_nvenc_* API.

The product is the C-ABI shared library ``libjmcodec_b200.so`` (sources in ``csrc/``, headers in
``/include``); this package is a thin ctypes binding used by the tests, ``bench.py`` and as the
example of how a host program binds it.  There is no CPU fallback: if the library has not been
built, importing :mod:`jmcodec_b200.lib` raises, and without a CUDA device every call fails.
"""
from .lib import (  # noqa: F401
    JMC_OP, JmcError, Ctx, Job, Frames, Pipeline, Event, NvDec, NvEnc, NvEncParam, RawPacket,
    device_count, last_error, lib_path, load, build, version, reload_env,
    JOB_ALIGNED16, JOB_LIST_ON_HOST, INLINE_LIST_MAX,
)
