"""Wall clock of compress2() with pinned host buffers (bench.py's e2e loop on its own): level-6 silesia-small.tar."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import zlib_rs_b200 as Z
from corpus import silesia_tar
d = silesia_tar(); N = len(d)
host = torch.frombuffer(bytearray(d), dtype=torch.uint8)
pin_in = torch.empty(N, dtype=torch.uint8).pin_memory(); pin_in.copy_(host)
L = Z.lib(); bound = int(L.compressBound(N))
pin_out = torch.empty(bound, dtype=torch.uint8).pin_memory()
ts = []
for i in range(15):
    n = ctypes.c_ulong(bound)
    t1 = time.perf_counter()
    rc = L.compress2(pin_out.data_ptr(), ctypes.byref(n), pin_in.data_ptr(), N, 6)
    ts.append(time.perf_counter() - t1)
    assert rc == 0
ts = sorted(ts[3:])
print("e2e compress2 ms: mean %.3f best %.3f  out %d  ZB_UPLOAD_CHUNKED=%s" % (sum(ts) / len(ts) * 1e3, ts[0] * 1e3, n.value, os.environ.get("ZB_UPLOAD_CHUNKED")))
