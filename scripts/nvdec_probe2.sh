#!/usr/bin/env bash
# second look at the NVDEC blocker: which libnvcuvid gets loaded, driver capabilities of the container, decoder engines
echo "NVIDIA_DRIVER_CAPABILITIES=${NVIDIA_DRIVER_CAPABILITIES:-<unset>}"
ldconfig -p | grep -i -E "nvcuvid|nvidia-encode" || echo "ldconfig: no nvcuvid / nvidia-encode"
find / -xdev -name "libnvcuvid*" 2>/dev/null | head
nvidia-smi -q | grep -i -E -A4 "^\s*(encoder|decoder|fbc) stats|Video" | head -30
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader | head -1
python - <<'PY'
import ctypes, os
try:
    lib = ctypes.CDLL("libnvcuvid.so.1")
    for line in open("/proc/self/maps"):
        if "nvcuvid" in line:
            path = line.split()[-1]; print("mapped:", path, os.path.getsize(path), "bytes"); break
except OSError as e:
    print("dlopen failed:", e)
PY
