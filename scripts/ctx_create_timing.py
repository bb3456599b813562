"""How long does CUDA context creation take per GPU, sequentially and from parallel threads? (diagnostic)"""
import ctypes as C, sys, threading, time
sys.path.insert(0, ".")
from commet_b200 import api
lib = api.load_library()
n = lib.commet_device_count()
mode = sys.argv[1]
hs = [C.c_void_p() for _ in range(n)]
t0 = time.perf_counter()
if mode == "seq":
    for g in range(n):
        t1 = time.perf_counter(); lib.commet_ctx_create(g, C.byref(hs[g])); print(f"  gpu {g}: {time.perf_counter()-t1:.3f} s")
else:
    th = [threading.Thread(target=lambda g=g: lib.commet_ctx_create(g, C.byref(hs[g]))) for g in range(n)]
    [t.start() for t in th]; [t.join() for t in th]
print(f"{mode}: {n} contexts in {time.perf_counter()-t0:.3f} s")
