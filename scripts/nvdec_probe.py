"""Time-boxed probe for SURVEY.md 8(f).3: is NVDEC usable from this image on the B200 box?
Looks for libnvcuvid (driver component, no headers in the image), dlopens it and asks
cuvidGetDecoderCaps for H.264 4:2:0 8-bit (struct layout restated from the public nvcuvid.h)."""
import ctypes
import ctypes.util
import glob
import json

import torch

out = {"libs": sorted(glob.glob("/usr/lib/x86_64-linux-gnu/libnvcuvid*") + glob.glob("/usr/lib64/libnvcuvid*") +
                      glob.glob("/usr/local/cuda/lib64/libnvcuvid*") + glob.glob("/usr/lib/x86_64-linux-gnu/libnvidia-encode*")),
       "find_library": ctypes.util.find_library("nvcuvid")}


class CUVIDDECODECAPS(ctypes.Structure):
    _fields_ = [("eCodecType", ctypes.c_int), ("eChromaFormat", ctypes.c_int), ("nBitDepthMinus8", ctypes.c_uint),
                ("reserved1", ctypes.c_uint * 3), ("bIsSupported", ctypes.c_ubyte), ("nNumNVDECs", ctypes.c_ubyte),
                ("nOutputFormatMask", ctypes.c_ushort), ("nMaxWidth", ctypes.c_uint), ("nMaxHeight", ctypes.c_uint),
                ("nMaxMBCount", ctypes.c_uint), ("nMinWidth", ctypes.c_ushort), ("nMinHeight", ctypes.c_ushort),
                ("bIsHistogramSupported", ctypes.c_ubyte), ("nCounterBitDepth", ctypes.c_ubyte),
                ("nMaxHistogramBins", ctypes.c_ushort), ("reserved3", ctypes.c_uint * 10)]


try:
    torch.cuda.init()
    torch.zeros(1, device="cuda")                       # primary context current on this thread
    lib = None
    for name in (out["libs"] + ["libnvcuvid.so.1", "libnvcuvid.so"]):
        try:
            lib = ctypes.CDLL(name)
            out["loaded"] = name
            break
        except OSError as e:
            out.setdefault("load_errors", []).append(f"{name}: {e}")
    if lib is not None:
        # make sure a driver-API context is current on this thread (torch's primary context normally is)
        cu = ctypes.CDLL("libcuda.so.1")
        ctx = ctypes.c_void_p()
        out["cuInit_rc"] = int(cu.cuInit(0))
        out["cuCtxGetCurrent_rc"] = int(cu.cuCtxGetCurrent(ctypes.byref(ctx)))
        out["ctx_current"] = bool(ctx.value)
        if not ctx.value:
            dev = ctypes.c_int()
            cu.cuDeviceGet(ctypes.byref(dev), 0)
            out["cuDevicePrimaryCtxRetain_rc"] = int(cu.cuDevicePrimaryCtxRetain(ctypes.byref(ctx), dev))
            out["cuCtxSetCurrent_rc"] = int(cu.cuCtxSetCurrent(ctx))
        caps = CUVIDDECODECAPS()
        caps.eCodecType, caps.eChromaFormat, caps.nBitDepthMinus8 = 4, 1, 0      # H.264, 4:2:0, 8 bit
        rc = lib.cuvidGetDecoderCaps(ctypes.byref(caps))
        out["cuvidGetDecoderCaps_rc"] = int(rc)
        out["h264"] = {"supported": int(caps.bIsSupported), "nvdecs": int(caps.nNumNVDECs), "max_w": int(caps.nMaxWidth),
                       "max_h": int(caps.nMaxHeight), "output_format_mask": int(caps.nOutputFormatMask)}
        out["has_parser"] = hasattr(lib, "cuvidCreateVideoParser")
except Exception as e:  # noqa: BLE001
    out["error"] = repr(e)
print(json.dumps(out))
