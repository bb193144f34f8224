import time, pynvml
pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0)
for fn in ("nvmlDeviceGetClockInfo",):
    t=time.perf_counter()
    for _ in range(20): pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    print(fn, (time.perf_counter()-t)/20*1e3, "ms")
g=getattr(pynvml,"nvmlDeviceGetCurrentClocksEventReasons",None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
t=time.perf_counter()
for _ in range(20): g(h)
print("reasons", (time.perf_counter()-t)/20*1e3, "ms")
