import time, pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
def t(f, n=20):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("clock sm        %.3f ms" % t(lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
print("max clock sm    %.3f ms" % t(lambda: nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
print("event reasons   %.3f ms" % t(lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
print("power           %.3f ms" % t(lambda: nv.nvmlDeviceGetPowerUsage(h)))
import sys; sys.path.insert(0, ".")
import torch, threading
from debwt_b200 import api
x = torch.empty(1 << 28, device="cuda")
def loop(stop, what):
    while not stop.is_set():
        what(); time.sleep(0.05)
def bench():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        x[:1 << 20].zero_(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 200 * 1e3
print("launch+sync baseline %.3f ms" % bench())
for name, f in (("clock", lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), ("reasons", lambda: nv.nvmlDeviceGetCurrentClocksEventReasons(h))):
    stop = threading.Event(); th = threading.Thread(target=loop, args=(stop, f)); th.start()
    print("launch+sync with %s polling %.3f ms" % (name, bench()))
    stop.set(); th.join()
