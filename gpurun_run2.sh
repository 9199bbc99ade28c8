python - <<'PY'
import time, traceback
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
    t=time.perf_counter()
    for i in range(5):
        print(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
    print("5 samples in", time.perf_counter()-t)
    print([n for n in dir(nv) if 'ThrottleReason' in n][:12])
except Exception:
    traceback.print_exc()
PY
